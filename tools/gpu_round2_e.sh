#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/step_timeline.py 1000000 60 > gpurun_out/e_timeline_heap.txt 2>&1
P2W_KNN_WARP=1 timeout 600 python tools/step_timeline.py 1000000 60 > gpurun_out/e_timeline_warp.txt 2>&1
P2W_KNN_WARP=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_warp.json 2>/dev/null
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/e_bench_heap.json 2>/dev/null
head -3 gpurun_out/e_timeline_heap.txt; grep -E "grid_query|query_|scan32|grid_" gpurun_out/e_timeline_heap.txt | head -20; echo; grep -E "grid_query|query_|scan32|grid_" gpurun_out/e_timeline_warp.txt | head -20
cut -c1-200 gpurun_out/e_bench_warp.json; cut -c1-200 gpurun_out/e_bench_heap.json
