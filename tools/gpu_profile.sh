#!/bin/bash
# round-2 profiling artefacts: launch list of the bench command, ncu --set full of the fused conv (three layers) and of the searches
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/p_bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 3 -c 3 -o gpurun_out/prof_conv_r2 python tools/prof_model.py 1000000 > gpurun_out/p_ncu_conv.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:grid_query -s 7 -c 7 -o gpurun_out/prof_knn_r2 python tools/prof_model.py 1000000 > gpurun_out/p_ncu_knn.log 2>&1
ls -la gpurun_out/*r2*; tail -2 gpurun_out/p_ncu_conv.log
