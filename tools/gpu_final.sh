#!/bin/bash
# Final 1-GPU evidence of a session: gpu_round (tests, ncu captures, launch list, bench), counters digested on the
# box, then the bench line that carries them.
tag=${1:-r02d}
bash tools/gpu_round.sh $tag
python tools/ncu_counters.py $tag 02 > gpurun_out/ncu_counters_$tag.log 2>&1
(timeout 900 python bench.py) > gpurun_out/bench_final_$tag.json 2> gpurun_out/bench_final_$tag.err
tail -c 600 gpurun_out/bench_final_$tag.json; tail -3 gpurun_out/bench_final_$tag.err
