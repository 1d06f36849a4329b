#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/parity_bf16.py > gpurun_out/i_parity_bf16.jsonl 2> gpurun_out/i_parity.err
timeout 600 python -m pytest tests/test_gpu_train.py -x -q 2>&1 | tail -5 > gpurun_out/i_pytest.txt
timeout 900 python bench.py --points 100000000 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/i_bench_100M_n1.json 2> gpurun_out/i_bench_100M_n1.err
cat gpurun_out/i_parity_bf16.jsonl; tail -3 gpurun_out/i_parity.err; tail -3 gpurun_out/i_pytest.txt; cut -c1-400 gpurun_out/i_bench_100M_n1.json; tail -3 gpurun_out/i_bench_100M_n1.err
