#!/bin/bash
# round-2 closing run: GPU suite, smoke, both bench arms, the dense micro-benchmark, launch list + ncu of the fused expand kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/f_pytest.txt
timeout 300 python __graft_entry__.py --smoke > gpurun_out/f_smoke.txt 2>&1
timeout 600 python bench.py > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err
timeout 300 python tools/bench_dense.py > gpurun_out/f_dense.jsonl 2> gpurun_out/f_dense.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dense_tc_kernel -c 2 -o gpurun_out/prof_dense_r2 python tools/prof_model.py 1000000 > gpurun_out/f_ncu_dense.log 2>&1
tail -4 gpurun_out/f_pytest.txt; tail -2 gpurun_out/f_smoke.txt; cut -c1-300 gpurun_out/f_bench1.json; cut -c1-300 gpurun_out/f_bench_ref.json; tail -2 gpurun_out/f_ncu_dense.log
