#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_distributed.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -5 > gpurun_out/q_pytest.txt
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/q_bench1.json 2>/dev/null
timeout 600 python tools/step_timeline.py 1000000 8 > gpurun_out/q_timeline.txt 2>&1
tail -3 gpurun_out/q_pytest.txt; cut -c1-220 gpurun_out/q_bench1.json; grep -E "unprofiled|ground_min|idle gaps" gpurun_out/q_timeline.txt
