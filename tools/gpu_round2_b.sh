#!/bin/bash
# round-2 GPU check B: heap kNN kernel parity + A/B, bench, sharded phases
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_fullsize.py tests/test_gpu_vote.py tests/test_gpu_engine.py -x -q 2>&1 | tail -15 > gpurun_out/b_pytest.txt
P2W_KNN_WARP=1 timeout 300 python tools/bench_knn.py > gpurun_out/b_knn_warp.jsonl 2>&1
timeout 300 python tools/bench_knn.py > gpurun_out/b_knn_heap.jsonl 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/b_bench1.json 2> gpurun_out/b_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_distributed_plot.py 4000000 --check > gpurun_out/b_dist2_4M.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/run_distributed_plot.py 2000000 > gpurun_out/b_dist2_2M.txt 2>&1
tail -4 gpurun_out/b_pytest.txt; cat gpurun_out/b_knn_warp.jsonl gpurun_out/b_knn_heap.jsonl | cut -c1-400; cat gpurun_out/b_bench1.json | cut -c1-300; tail -1 gpurun_out/b_dist2_4M.txt; tail -1 gpurun_out/b_dist2_2M.txt
