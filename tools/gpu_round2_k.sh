#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_engine.py tests/test_gpu_model.py -x -q 2>&1 | tail -6 > gpurun_out/k_pytest.txt
P2W_CONV_G2=0 timeout 600 python tools/step_timeline.py 1000000 12 > gpurun_out/k_timeline_g2off.txt 2>&1
timeout 600 python tools/step_timeline.py 1000000 12 > gpurun_out/k_timeline_g2on.txt 2>&1
P2W_CONV_G2=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_g2off.json 2>/dev/null
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/k_bench_g2on.json 2>/dev/null
tail -3 gpurun_out/k_pytest.txt; grep -E "unprofiled|conv_tc" gpurun_out/k_timeline_g2off.txt; grep -E "unprofiled|conv_tc" gpurun_out/k_timeline_g2on.txt
python - <<'PY'
import json
for f in ("gpurun_out/k_bench_g2off.json","gpurun_out/k_bench_g2on.json"):
    d=json.loads([l for l in open(f) if l.startswith("{")][0]); print(f, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["avg_launch_ms"], d["roofline"]["frac"])
PY
