#!/bin/bash
# Same-box A/B of K2 variants with the tuning library (FASTB_TUNE=1 python build_fastb.py).
# usage: tools/tune_sweep.sh <out file> ; each line: workload, env, value, frac
out=${1:-gpurun_out/tune_sweep.txt}
export FASTB_LIBRARY=$PWD/fast_b200/libfastb_tune.so
run() {  # workload, env assignments...
  w=$1; shift
  line=$(env "$@" python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>/dev/null | tail -1)
  python - "$w" "$*" <<PY >> $out
import json, sys
try:
    d = json.loads('''$line''')
    print(f"{sys.argv[1]:3s} {sys.argv[2]:40s} value {d['value']/1e6:8.4f} M/s  frac {d['roofline']['frac']:.4f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  mean_r {d['check']['mean_r']:.6f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e)
PY
}
: > $out
for w in c4 c5; do
  run $w X=0
  run $w FASTB_SHFL=0
done
run c4 FASTB_SHAPE=256x3
run c4 FASTB_SHAPE=512x1
run c4 FASTB_SHAPE=128x4
run c4 FASTB_SHAPE=128x6
run c4 FASTB_STAGE=0
run c4 FASTB_SHAPE=256x3 FASTB_STAGE=0
run c5 FASTB_SHAPE=256x2
run c5 FASTB_SHAPE=256x4
run c5 FASTB_SHAPE=128x4
run c5 FASTB_SHAPE=512x1
run c5 FASTB_STAGE=1
run c2 X=0
run c2 FASTB_SHAPE=128x5
run c2 FASTB_SHAPE=128x6
run c2 FASTB_SHAPE=256x2
cat $out
