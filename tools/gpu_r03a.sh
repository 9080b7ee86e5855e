#!/bin/bash
# GPU session r03a: chirp-z kernel v5 -- tests, two-copy vs one-copy FFT, ncu capture, quick bench of the radix workloads
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -25) > gpurun_out/pytest_r03a.log 2>&1
tail -6 gpurun_out/pytest_r03a.log
echo "--- two-copy (product)"; timeout 120 python tools/run_c1prime.py 2>&1 | tail -2 | tee gpurun_out/c1prime_two_r03a.txt
echo "--- one-copy (tune)"; FASTB_LIBRARY=$PWD/fast_b200/libfastb_tune_one.so timeout 120 python tools/run_c1prime.py 2>&1 | tail -2 | tee gpurun_out/c1prime_one_r03a.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:screen_detect_bluestein -s 1 -c 1 \
    -f -o gpurun_out/prof_c1prime_r03a python tools/run_c1prime.py > gpurun_out/ncu_c1prime_r03a.log 2>&1
bash tools/quick_bench.sh gpurun_out/quick_bench_r03a.txt
