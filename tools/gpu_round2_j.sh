#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sanitize_case.py > gpurun_out/j_plain.txt 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 1500 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_case.py > gpurun_out/j_sanitizer_$tool.txt 2>&1
  tail -4 gpurun_out/j_sanitizer_$tool.txt
done
tail -2 gpurun_out/j_plain.txt
