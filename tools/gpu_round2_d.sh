#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_query_heap -s 3 -c 1 -o gpurun_out/d_heap_uniform32 python tools/bench_knn.py uniform32 > gpurun_out/d_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:grid_query_heap -s 3 -c 1 -o gpurun_out/d_heap_vote python tools/bench_knn.py vote > gpurun_out/d_ncu2.log 2>&1
ls -la gpurun_out/*.ncu-rep; tail -3 gpurun_out/d_ncu1.log
