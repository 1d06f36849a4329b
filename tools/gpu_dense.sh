#!/bin/bash
# the fused expand kernel: parity, micro-benchmark, A/B in the step
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_engine.py -x -q -k "dense_expand" > gpurun_out/d_pytest.txt 2>&1; tail -5 gpurun_out/d_pytest.txt
timeout 300 python tools/bench_dense.py > gpurun_out/d_dense.jsonl 2>gpurun_out/d_dense.err; grep -E "expand" gpurun_out/d_dense.jsonl
P2W_DENSE_TC=1 timeout 300 python -m pytest tests/test_gpu_engine.py tests/test_gpu_fullsize.py -x -q > gpurun_out/d_pytest_on.txt 2>&1; tail -3 gpurun_out/d_pytest_on.txt
for i in 1 2; do
P2W_DENSE_TC=0 timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('off', d['ms_per_step'])"
P2W_DENSE_TC=1 timeout 300 python bench.py --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('on ', d['ms_per_step'])"
done
