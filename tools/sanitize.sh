#!/bin/bash
# compute-sanitizer (memcheck + racecheck) over small cases of every kernel family; summaries under gpurun_out/sanitizer_*.txt
# usage (on the GPU box): bash tools/sanitize.sh
out=gpurun_out
mkdir -p $out
for tool in memcheck racecheck; do
  log=$out/sanitizer_$tool.txt
  : > $log
  for m in 0 1 2 3 4 5 6 7 8 9 10 11; do
    echo "=== $tool: mode $m (600 bp graph, 6 paths, 4 reads x 120 bp)" >> $log
    timeout 300 compute-sanitizer --tool $tool python tools/small_run.py $m 4 2>&1 | grep -E "^mode|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | head -8 >> $log
  done
  echo "=== $tool: mode 2, 1 400-base reads (striped kernel), 900 bp graph" >> $log
  timeout 300 compute-sanitizer --tool $tool python tools/small_run.py 2 3 900 3 1400 2>&1 | grep -E "^mode|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | head -8 >> $log
  echo "=== $tool: mode 9, 300-base reads, 1 500 bp graph, 40 paths" >> $log
  timeout 300 compute-sanitizer --tool $tool python tools/small_run.py 9 3 1500 40 300 2>&1 | grep -E "^mode|ERROR SUMMARY|RACECHECK SUMMARY|Invalid|hazard" | head -8 >> $log
done
tail -n 100 $out/sanitizer_memcheck.txt $out/sanitizer_racecheck.txt
