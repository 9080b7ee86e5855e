#!/bin/bash
out=${1:-gpurun_out/tune_sweep5.txt}
export FASTB_LIBRARY=$PWD/fast_b200/libfastb_tune.so
run() {
  w=$1; shift
  line=$(env "$@" python bench.py --workload $w --steps 5 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>/dev/null | tail -1)
  python - "$w" "$*" <<PY >> $out
import json, sys
try:
    d = json.loads('''$line''')
    print(f"{sys.argv[1]:3s} {sys.argv[2]:40s} value {d['value']/1e6:8.4f} M/s  frac {d['roofline']['frac']:.4f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  e2e {d['e2e']['value']/1e6:8.4f}  mean_r {d['check']['mean_r']:.6f}")
except Exception as e:
    print(sys.argv[1], sys.argv[2], 'FAILED', e)
PY
}
: > $out
run c2 X=0
for s in 600 1200 2400 5000; do run c2 FASTB_STAGGER=$s; done
run c4 X=0
for s in 600 1500 3000; do run c4 FASTB_WSTAGGER=$s; done
run c5 X=0
for s in 600 1500 3000; do run c5 FASTB_WSTAGGER=$s; done
cat $out
