#!/bin/bash
# Same-box A/B of K2 variants with the tuning library (FASTB_TUNE=1 python build_fastb.py -> libfastb_tune.so).
#   tools/tune_sweep.sh <out file> "<workload> [ENV=VALUE ...]" ...
# e.g. tools/tune_sweep.sh gpurun_out/t.txt "c4" "c4 FASTB_SHFL=0" "c5 FASTB_SHAPE=256x3" "c2 FASTB_ONCHIP=384"
# Switches: fast_b200/csrc/tune/screen_detect_tune.cu.  Each line: workload, switches, value, model-A fraction,
# kernel ms, e2e, check.mean_r (equal across variants of one stream = identical results).
out=$1; shift
export FASTB_LIBRARY=$PWD/fast_b200/libfastb_tune.so
: > $out
for spec in "$@"; do
  set -- $spec
  w=$1; shift
  line=$(env "$@" X=0 python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>/dev/null | tail -1)
  python - "$w" "$*" <<PY >> $out
import json, sys
try:
    d = json.loads('''$line''')
    print(f"{sys.argv[1]:3s} {sys.argv[2]:40s} value {d['value']/1e6:8.4f} M/s  frac {d['roofline']['frac']:.4f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  e2e {d['e2e']['value']/1e6:8.4f}  mean_r {d['check']['mean_r']:.6f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e)
PY
done
cat $out
