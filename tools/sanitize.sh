#!/bin/bash
# compute-sanitizer over every kernel family -> gpurun_out/sanitizer_<tag>.txt
tag=${1:-r02}
out=gpurun_out/sanitizer_$tag.txt
echo "# compute-sanitizer over every kernel family (tests/sanitize_workload.py); B200, CUDA 12.9" > $out
for tool in memcheck racecheck initcheck synccheck; do
  echo "=== compute-sanitizer --tool $tool ===" >> $out
  timeout 900 compute-sanitizer --tool $tool python tests/sanitize_workload.py 2>&1 | grep -E "COMPUTE-SANITIZER|sanitize workload|SUMMARY|Error|error|hazard|Hazard" | head -40 >> $out
done
cat $out
