#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_vote.py tests/test_gpu_distributed.py -x -q 2>&1 | tail -15 > gpurun_out/c_pytest.txt
timeout 300 python tools/bench_knn.py > gpurun_out/c_knn_heap.jsonl 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/c_bench1.json 2> gpurun_out/c_bench1.err
tail -4 gpurun_out/c_pytest.txt; cut -c1-330 gpurun_out/c_knn_heap.jsonl; cut -c1-300 gpurun_out/c_bench1.json; tail -3 gpurun_out/c_bench1.err
