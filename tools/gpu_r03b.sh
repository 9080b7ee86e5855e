#!/bin/bash
# GPU session r03b: full tests after the M = 512 shuffle stage / one-launch TEMPORAL run, chirp-z size sweep, C1 timing
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/pytest_r03b.log 2>&1
tail -4 gpurun_out/pytest_r03b.log
timeout 300 python tools/chirpz_sizes.py 2>&1 | tee gpurun_out/chirpz_sizes_r03b.txt
timeout 300 python - <<'PY' 2>&1 | tee gpurun_out/c1_temporal_r03b.txt
import time, torch, fast_b200
from fast_b200 import configs
sim = fast_b200.Fast(configs.c1(seed=1))
for k in range(3): sim.run()
torch.cuda.synchronize(); t0 = time.perf_counter()
for k in range(20): sim.run()
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print('C1 TEMPORAL verbatim: %.3f ms per run, %.1f K steps/s' % (1e3 * dt, sim.Niter / dt / 1e3))
PY
