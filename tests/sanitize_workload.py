"""Small invocations of every kernel family for compute-sanitizer:
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python tests/sanitize_workload.py
(profiles/sanitizer_r02.txt holds the round-2 output).  Not collected by pytest."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import fast_b200
from fast_b200 import configs, comms, dist, sweep, _lib

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
# round 2: the fast noise stream, chirp-z on 164 and on a 1000-point grid (M = 2048, named line barriers), fused
# statistics, a batched sweep with per-item statistics, layer screens through radix / chirp-z / direct
for name, n, kw in (('c1prime', 8, {'RNG': 'device-fast'}), ('c2', 16, {'RNG': 'device-fast'}), ('c4', 8, {'RNG': 'device-fast'}),
                    ('c1prime', 8, {'RNG': 'numpy'}), ('c4', 4, {'RNG': 'numpy'})):
    p = getattr(configs, name)(niter=n, nchunks=1)
    p.update(kw, SEED=2)
    sim = fast_b200.Fast(p)
    sb = dist.StatsBuffers(64, sim.device)
    if kw['RNG'] != 'numpy':
        sim.screen_detect(0, n // 2, stats=sb)
        assert sb.summary()['n'] == n
    assert np.isfinite(np.abs(sim.run()._r)).all()
ps = [configs.c3_elevation(e, niter=8, nchunks=2, seed=5 + i) for i, e in enumerate((15.0, 50.0, 80.0))]
res = sweep.run_sweep(sweep.build_sims(ps), stats=True, nbins=32)
assert all(np.isfinite(r._r).all() for r in res)
for N in (128, 164, 1000, 1100):
    w = _lib.make_weight(torch.rand(2, N, N, dtype=torch.float64, device='cuda') * 1e-5, 1.0)
    assert torch.isfinite(_lib.layer_screens(w, 3)).all()
N, P = 1000, 200
rp = _lib.RunParams()
rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.seed, rp.algo = N, P, 400, 2, 2, 17, _lib.ALGO_AUTO
rp.u_sum, rp.sigma_chi = float(P * P), 0.02
wt = _lib.make_weight(torch.rand(N, N, dtype=torch.float64, device='cuda') * 1e-6, 1.0)
ws = torch.empty(_lib.screen_detect_workspace_bytes(rp), dtype=torch.uint8, device='cuda')
oa = torch.empty(2, dtype=torch.float32, device='cuda'); ob = torch.empty_like(oa)
_lib.screen_detect(rp, wt, torch.ones(P, P, dtype=torch.float32, device='cuda'), oa, ob, ws)
assert torch.isfinite(oa).all()
# the chirp-z kernels over transform lengths and cell classes (M = 64 .. 1024; M = 512 runs the shuffle last stage),
# device RNG and caller-supplied noise
for N, lo, P in ((58, 26, 6), (104, 40, 24), (236, 108, 20), (200, 78, 44), (300, 60, 180), (460, 210, 50), (700, 250, 200)):
    rp = _lib.RunParams()
    rp.n, rp.n_pup, rp.lo, rp.n_pairs, rp.pairs_per_chunk, rp.seed, rp.algo = N, P, lo, 3, 3, 17, _lib.ALGO_AUTO
    rp.u_sum, rp.sigma_chi = float(P * P), 0.02
    wt = _lib.make_weight(torch.rand(N, N, dtype=torch.float64, device='cuda') * 1e-6, 1.0)
    ws = torch.empty(_lib.screen_detect_workspace_bytes(rp), dtype=torch.uint8, device='cuda')
    oa = torch.empty(3, dtype=torch.float32, device='cuda'); ob = torch.empty_like(oa)
    U1 = torch.ones(P, P, dtype=torch.float32, device='cuda')
    _lib.screen_detect(rp, wt, U1, oa, ob, ws)
    assert torch.isfinite(oa).all()
    if N <= 300:
        nz = torch.randn(3, N, N, 2, dtype=torch.float32, device='cuda')
        _lib.screen_detect(rp, wt, U1, oa, ob, ws, chi=torch.zeros(6, dtype=torch.float32, device='cuda'), noise=nz)
        assert torch.isfinite(ob).all()
x = np.exp(0.3 * np.random.default_rng(0).standard_normal(5000)).astype(np.float32)
comms.ber_ook(np.array([3.0, 9.0]), x); comms.sep_qam(16, 10.0, x)
comms.fade_prob(x, np.array([0.5, 0.9])); comms.fade_dur(x, 0.8)
m = comms.Modulator(x[:500], '16-QAM', EsN0=10.0, symbols_per_iter=20); m.run(); m.modulate(); m.demodulate()
comms.convolve_awgn_qam(np.sqrt(x), 16, 24, 12.0); comms.convolve_awgn_qam(np.sqrt(x), 4, 24, 12.0, region_size='full', shot=True)
comms.generalised_mutual_information_qam(np.sqrt(x), 16, 24, 12.0)
torch.cuda.synchronize()
print('sanitize workload done, launches', _lib.launch_count())
