#!/bin/bash
# Experiment 15 (profiles/experiments_r02.txt): pass-2 column prefetch through L1 and more resident warps against the
# product kernel.  Needs the tuning library:  FASTB_TUNE=1 FASTB_TUNE_TAG=_pf FASTB_PF=1 python build_fastb.py
mkdir -p gpurun_out
one() {  # label, workload, env...
  label=$1; w=$2; shift 2
  line=$(env "$@" X=0 python bench.py --workload $w --steps 6 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>/dev/null | tail -1)
  python - "$label" "$w" <<PY
import json, sys
try:
    d = json.loads('''$line''')
    print(f"{sys.argv[1]:28s} {sys.argv[2]:3s} value {d['value']/1e6:8.4f} M/s  frac {d['roofline']['frac']:.4f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  mean_r {d['check']['mean_r']:.6f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e)
PY
}
PF=$PWD/fast_b200/libfastb_tune_pf.so
{
for w in c2 c4 c5; do one product $w; done
for w in c2 c4 c5; do one prefetch $w FASTB_LIBRARY=$PF; done
one "prefetch 256x3" c2 FASTB_LIBRARY=$PF FASTB_SHAPE=256x3
one "prefetch 256x2" c2 FASTB_LIBRARY=$PF FASTB_SHAPE=256x2
one "prefetch 128x5" c2 FASTB_LIBRARY=$PF FASTB_SHAPE=128x5
one product c2
} | tee gpurun_out/tune_r03c_prefetch.txt
(FASTB_LIBRARY=$PF timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "radix or window or device_rng_run or reference_noise or launch_split" 2>&1 | tail -4) | tee gpurun_out/pytest_r03c_pf.log
