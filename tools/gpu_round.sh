#!/bin/bash
# One GPU session of the round: tests, bench, launch list, ncu captures (tag = $1).
tag=${1:-r02a}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/pytest_$tag.log 2>&1
tail -4 gpurun_out/pytest_$tag.log
B="python bench.py --steps 1 --warmup 3 --no-cpu --no-comparator --no-per-workload"
for w in c2 c4 c5; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_detect_radix -s 4 -c 1 \
      -f -o gpurun_out/prof_${w}_$tag $B --workload $w > gpurun_out/ncu_${w}_$tag.log 2>&1
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_detect_bluestein -s 1 -c 1 \
    -f -o gpurun_out/prof_c1prime_$tag python tools/run_c1prime.py > gpurun_out/ncu_c1prime_$tag.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-comparator > gpurun_out/launches_$tag.log 2>&1
(timeout 600 python bench.py) > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
head -c 1500 gpurun_out/bench_$tag.json; tail -3 gpurun_out/bench_$tag.err
