#!/bin/bash
out=${1:-gpurun_out/tune_sweep3.txt}
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
for w in c2 c4 c5; do
  run $w X=0
  run $w FASTB_L2PERSIST=1
done
cat $out
for w in c2 c4 c5; do
FASTB_L2PERSIST=1 timeout 200 ncu --metrics dram__bytes_write.sum,dram__bytes_read.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum --clock-control none -k regex:screen_detect_radix -s 4 -c 1 python bench.py --workload $w --steps 1 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>&1 | grep -E "dram__|lts__|gpu__time" | sed "s/^/$w persist: /" | tee -a $out
done
