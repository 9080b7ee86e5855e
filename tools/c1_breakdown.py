import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time, torch, numpy as np, fast_b200
from fast_b200 import configs, temporal, funcs
sim = fast_b200.Fast(configs.c1(seed=1))
for k in range(3): sim.run()
def T(f, n=30):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print('run total        %.3f ms' % T(sim.run))
print('compute_logamp   %.3f ms' % T(sim.compute_logamp))
print('phs_temporal(0)  %.3f ms' % T(lambda: sim.compute_phs_temporal(chunk=0)))
print('layer_screens    %.3f ms' % T(lambda: fast_b200._lib.layer_screens(sim._d['weight_per_layer'], 5)))
step = sim.pixel_shifts[:, :, -1, None, None]
st = np.stack([sim.interp_coords + k * step for k in range(sim.Nchunks)])
print('sample_coords    %.3f ms' % T(lambda: temporal.sample_coordinates(st, sim.Npxls)))
print('_run_temporal    %.3f ms' % T(sim._run_temporal))
flat = sim._run_temporal()
print('_publish         %.3f ms' % T(lambda: sim._publish(flat)))
