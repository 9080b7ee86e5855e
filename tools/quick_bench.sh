#!/bin/bash
# value / frac / e2e of the three single-GPU workloads with the product library
out=${1:-gpurun_out/quick_bench.txt}
: > $out
for w in c2 c4 c5; do
  line=$(python bench.py --workload $w --steps 8 --warmup 3 --no-cpu --no-comparator --no-per-workload 2>/dev/null | tail -1)
  python - "$w" <<PY >> $out
import json, sys
try:
    d = json.loads('''$line''')
    print(f"{sys.argv[1]:3s} value {d['value']/1e6:8.4f} M/s  frac {d['roofline']['frac']:.4f}  kernel_ms {d['roofline']['kernel_ms']:.3f}  e2e {d['e2e']['value']/1e6:8.4f}  mean_r {d['check']['mean_r']:.6f}")
except Exception as e:
    print(sys.argv[1], 'FAILED', e)
PY
done
cat $out
