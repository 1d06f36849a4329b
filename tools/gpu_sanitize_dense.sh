#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/sanitize_dense.txt
for tool in memcheck racecheck synccheck; do
  echo "## --tool $tool (tools/sanitize_case.py dense)" >> gpurun_out/sanitize_dense.txt
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_case.py dense 2>&1 | grep -E "sanitize_case|SUMMARY|COMPUTE-SANITIZER|Error|error|hazard" | head -20 >> gpurun_out/sanitize_dense.txt
done
cat gpurun_out/sanitize_dense.txt
timeout 600 python bench.py > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; python -c "
import json; d=json.load(open('gpurun_out/f_bench1.json')); print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'])"
