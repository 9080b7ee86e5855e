#!/bin/bash
out=${1:-gpurun_out/tune_sweep4.txt}
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
run c2 FASTB_ONCHIP=384
run c2 FASTB_ONCHIP=256
run c2 FASTB_SHAPE=256x2
cat $out
FASTB_ONCHIP=384 timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_detect_radix -s 4 -c 1 -f -o gpurun_out/prof_c2onchip_r02b python bench.py --workload c2 --steps 1 --warmup 3 --no-cpu --no-comparator --no-per-workload > gpurun_out/ncu_c2onchip_r02b.log 2>&1
tail -2 gpurun_out/ncu_c2onchip_r02b.log
