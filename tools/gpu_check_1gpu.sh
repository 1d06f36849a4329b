#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/l_pytest.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/l_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/l_bench1.json 2> gpurun_out/l_bench1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/l_bench_ref.json 2> gpurun_out/l_bench_ref.err
tail -4 gpurun_out/l_pytest.txt; tail -2 gpurun_out/l_smoke.txt; cut -c1-300 gpurun_out/l_bench1.json; cut -c1-400 gpurun_out/l_bench_ref.json
