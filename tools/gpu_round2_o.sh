#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/run_distributed_plot.py 2000000 > gpurun_out/o_dist2_2M.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus 2 --steps 8 --warmup 3 > gpurun_out/o_bench2.json 2> gpurun_out/o_bench2.err
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/o_bench1.json 2>/dev/null
tail -1 gpurun_out/o_dist2_2M.txt| cut -c1-700; grep '^{' gpurun_out/o_bench2.json | cut -c1-220; cut -c1-220 gpurun_out/o_bench1.json
