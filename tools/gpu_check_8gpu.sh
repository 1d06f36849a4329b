#!/bin/bash
# round-2 eight-GPU run: BASELINE.json configs[3] (one 100 M-point plot sharded over 8 B200), the default weak series, configs[4]
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/g_gpus.txt
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 200)) "$@"; }
run 8 tools/run_distributed_plot.py 16000000 --check > gpurun_out/g_check_16M_n8.txt 2>&1
run 8 bench.py --gpus 8 --scaling strong --points 100000000 --steps 3 --warmup 3 > gpurun_out/g_bench_100M_n8.json 2> gpurun_out/g_bench_100M_n8.err
run 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/g_bench_weak_n8.json 2> gpurun_out/g_bench_weak_n8.err
run 4 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/g_bench_weak_n4.json 2> gpurun_out/g_bench_weak_n4.err
run 8 tools/run_distributed_plot.py 100000000 --steps 1 > gpurun_out/g_phases_100M_n8.txt 2>&1
run 8 tools/bench_train.py > gpurun_out/g_train_n8.txt 2>&1
run 4 tools/bench_train.py > gpurun_out/g_train_n4.txt 2>&1
tail -1 gpurun_out/g_check_16M_n8.txt | cut -c1-1500; grep "^{" gpurun_out/g_bench_100M_n8.json | cut -c1-300; tail -2 gpurun_out/g_bench_100M_n8.err; grep "^{" gpurun_out/g_bench_weak_n8.json | cut -c1-260; grep "^{" gpurun_out/g_bench_weak_n4.json | cut -c1-260; tail -1 gpurun_out/g_phases_100M_n8.txt | cut -c1-700; grep "^{" gpurun_out/g_train_n8.txt; grep "^{" gpurun_out/g_train_n4.txt
