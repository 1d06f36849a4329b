#!/bin/bash
# round-2 GPU check A: tests, 1-GPU bench, sharded plot on 2 GPUs (check + bench)
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/a_pytest.txt
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/a_bench1.json 2> gpurun_out/a_bench1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/run_distributed_plot.py 4000000 --check > gpurun_out/a_dist2_check.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/a_bench2.json 2> gpurun_out/a_bench2.err
tail -5 gpurun_out/a_pytest.txt; cat gpurun_out/a_bench1.json; tail -3 gpurun_out/a_dist2_check.txt; cat gpurun_out/a_bench2.json; tail -5 gpurun_out/a_bench2.err
