#!/bin/bash
# host-to-host step time: default run, without the CPU leg, with the fused expand off
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; grep "host-to-host" gpurun_out/f_bench1.err
python -c "
import json; d=json.load(open('gpurun_out/f_bench1.json')); print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 300 python bench.py --no-cpu-baseline 2>&1 >/dev/null | grep "host-to-host"
P2W_DENSE_TC=0 timeout 300 python bench.py --no-cpu-baseline 2>&1 >/dev/null | grep "host-to-host"
timeout 300 python bench.py --no-cpu-baseline --steps 8 2>&1 >/dev/null | grep "host-to-host"
