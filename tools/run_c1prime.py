"""One warm-up and two timed launches of the C1' workload (N = 164, chirp-z kernel) -- ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fast_b200
from fast_b200 import configs
n_real = 200000
sim = fast_b200.Fast(configs.c1prime(niter=n_real, nchunks=1, seed=1))
sim.screen_detect(0, 2000)
for r in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.screen_detect((r + 1) * 100000, 100000); e1.record(); e1.synchronize()
    print('c1prime launch %d: %.3f ms -> %.3f M realisations/s' % (r, e0.elapsed_time(e1), n_real / e0.elapsed_time(e1) / 1e3))
