#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_distributed.py tests/test_gpu_pipeline.py -x -q 2>&1 | tail -8 > gpurun_out/h_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_distributed_plot.py 4000000 --check > gpurun_out/h_dist2_4M.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/h_bench2.json 2> gpurun_out/h_bench2.err
tail -3 gpurun_out/h_pytest.txt; tail -1 gpurun_out/h_dist2_4M.txt | cut -c1-900; grep '^{' gpurun_out/h_bench2.json | cut -c1-250
