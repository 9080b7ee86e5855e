"""Ad-hoc timing of the pieces of sweep.run_sweep (C3) on one GPU."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import fast_b200
from fast_b200 import configs, sweep, dist, _lib

ps = [configs.c3_elevation(e, niter=10000, nchunks=1, seed=100 + i) for i, e in enumerate(configs.C3_ELEVATIONS)]
sims = sweep.build_sims(ps)
sweep.run_sweep(sims)
torch.cuda.synchronize()
for stats in (False, True):
    t0 = time.perf_counter()
    sweep.run_sweep(sims, stats=stats)
    torch.cuda.synchronize()
    print('run_sweep stats=%s wall %.2f ms' % (stats, 1e3 * (time.perf_counter() - t0)))
# pieces
t0 = time.perf_counter()
keys = [sweep._group_key(s) for s in sims]
print('group keys %.2f ms' % (1e3 * (time.perf_counter() - t0)))
lead = sims[0]
E, ppi = 16, 5000
weights = torch.stack([s._d['weight'] for s in sims]).contiguous()
sigma = torch.tensor([s.logamp_var ** 0.5 for s in sims], dtype=torch.float32, device='cuda')
import numpy
seeds = torch.from_numpy(numpy.array([s._run_seed() for s in sims], dtype=numpy.uint64).view(numpy.int64)).cuda()
rp = lead._run_params(E * ppi, 0)
ws = torch.empty(_lib.screen_detect_workspace_bytes(rp, E), dtype=torch.uint8, device='cuda')
a = torch.empty(E * ppi, dtype=torch.float32, device='cuda'); b = torch.empty_like(a)
for use_stats in (False, True):
    sb = dist.StatsBuffers(4096, 'cuda', n_items=E)
    st = _lib.run_stats(-60, 3, 4096, sb.sums, sb.minmax, sb.hist) if use_stats else None
    for rep in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        _lib.screen_detect_batch(rp, weights, lead._d['U'], a, b, ws, batch=dict(n_items=E, pairs_per_item=ppi, sigma_chi=sigma, seeds=seeds), stats=st)
        e1.record(); e1.synchronize()
        print('batch kernel stats=%s: %.2f ms' % (use_stats, e0.elapsed_time(e1)))
# single config same count
sim = fast_b200.Fast(configs.c2(niter=160000, nchunks=1, seed=1))
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.screen_detect(0, 80000); e1.record(); e1.synchronize()
    print('single config 80000 pairs: %.2f ms' % e0.elapsed_time(e1))
sb1 = dist.StatsBuffers(4096, 'cuda')
for rep in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); sim.screen_detect(0, 80000, stats=sb1); e1.record(); e1.synchronize()
    print('single config 80000 pairs + stats: %.2f ms' % e0.elapsed_time(e1))
