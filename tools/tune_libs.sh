#!/bin/bash
# A/B of whole tuning libraries (compile-time variants): tools/tune_libs.sh <out> <lib tag>...
out=$1; shift
: > $out
for tag in "$@"; do
  if [ "$tag" = product ]; then unset FASTB_LIBRARY; else export FASTB_LIBRARY=$PWD/fast_b200/libfastb_tune$tag.so; fi
  for w in c2 c4 c5; do
    line=$(python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>/dev/null | tail -1)
    python - "$w" "$tag" <<PY >> $out
import json, sys
try:
    d = json.loads('''$line''')
    print(f"{sys.argv[2]:10s} {sys.argv[1]:3s} value {d['value']/1e6:8.4f} M/s  frac {d['roofline']['frac']:.4f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  mean_r {d['check']['mean_r']:.6f}")
except Exception as e:
    print(sys.argv[2], sys.argv[1], 'FAILED', e)
PY
  done
done
cat $out
