#!/bin/bash
# Multi-GPU session (gpurun --gpus N): the sharded bit-identity tests, then the bench under torchrun
n=${1:-2}; tag=${2:-r02d}
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -q -k "sharded or device_key or multi" 2>&1 | tail -4) | tee gpurun_out/multi_gpu_check_${tag}_${n}gpu.log
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29655 \
   bench.py --gpus $n --steps 10 --warmup 3) > gpurun_out/bench_${tag}_${n}gpu.json 2> gpurun_out/bench_${tag}_${n}gpu.err
tail -c 900 gpurun_out/bench_${tag}_${n}gpu.json; tail -2 gpurun_out/bench_${tag}_${n}gpu.err
