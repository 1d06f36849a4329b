#!/bin/bash
# the round-end sequence: reference arm, then the GPU arm on the same box; per-step host-to-host times from stderr
mkdir -p gpurun_out
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; cut -c1-120 gpurun_out/f_bench_ref.json
timeout 600 python bench.py > gpurun_out/f_bench1.json 2> gpurun_out/f_bench1.err; grep "host-to-host" gpurun_out/f_bench1.err
python -c "
import json; d=json.load(open('gpurun_out/f_bench1.json')); print(d['ms_per_step'], d['e2e'], d['roofline']['frac'], d['cpu_baseline']['value'])"
timeout 200 python __graft_entry__.py --smoke 2>&1 | tail -1
timeout 600 python bench.py --no-cpu-baseline 2>&1 >/dev/null | grep "host-to-host"
