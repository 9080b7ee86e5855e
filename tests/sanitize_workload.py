"""Small invocations of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tests/sanitize_workload.py
(profiles/sanitizer_r01.txt holds the round-1 output: 0 errors, 0 hazards).  Not collected by pytest."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fast_b200
from fast_b200 import configs, comms, _lib

for kw in ({}, {'COHERENT': True}, {'SUBHARM': True}, {'AO_MODE': 'LGSAO'}, {'MODAL': True, 'ZMAX': 6},
           {'TEMPORAL': True, 'NITER': 20, 'NCHUNKS': 2, 'DT': 0.002}):
    p = configs.mini()
    p.update({'NITER': 16, 'NCHUNKS': 2, 'SEED': 3})
    p.update(kw)
    r = fast_b200.Fast(p).run()
    assert np.isfinite(np.abs(r._r)).all()
for name, n in (('c1prime', 8), ('c2', 64), ('c4', 8), ('c5', 4)):
    p = getattr(configs, name)(niter=n, nchunks=1)
    p['SEED'] = 1
    r = fast_b200.Fast(p).run()
    assert np.isfinite(np.abs(r._r)).all()
x = np.exp(0.3 * np.random.default_rng(0).standard_normal(5000)).astype(np.float32)
comms.ber_ook(np.array([3.0, 9.0]), x); comms.sep_qam(16, 10.0, x)
comms.fade_prob(x, np.array([0.5, 0.9])); comms.fade_dur(x, 0.8)
m = comms.Modulator(x[:500], '16-QAM', EsN0=10.0, symbols_per_iter=20); m.run(); m.modulate(); m.demodulate()
comms.convolve_awgn_qam(np.sqrt(x), 16, 24, 12.0); comms.convolve_awgn_qam(np.sqrt(x), 4, 24, 12.0, region_size='full', shot=True)
comms.generalised_mutual_information_qam(np.sqrt(x), 16, 24, 12.0)
torch.cuda.synchronize()
print('sanitize workload done, launches', _lib.launch_count())
