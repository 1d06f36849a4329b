#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_vote.py tests/test_gpu_ops.py -x -q 2>&1 | tail -4 > gpurun_out/r_pytest.txt
P2W_KNN_BINQ=0 timeout 300 python tools/bench_knn.py > gpurun_out/r_knn_binq0.jsonl 2>&1
P2W_KNN_BINQ=1 timeout 300 python tools/bench_knn.py > gpurun_out/r_knn_binq1.jsonl 2>&1
P2W_KNN_BINQ=1 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_binq1.json 2>/dev/null
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_default.json 2>/dev/null
P2W_KNN_BINQ=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r_bench_binq0.json 2>/dev/null
tail -2 gpurun_out/r_pytest.txt
python - <<'PY'
import json
for f in ("r_knn_binq0","r_knn_binq1"):
    print(f)
    for l in open(f"gpurun_out/{f}.jsonl"):
        try: d=json.loads(l)
        except Exception: continue
        print("  ", d["case"][:50], d["tiles"], d["k"], d["ms"])
for f in ("r_bench_binq0","r_bench_default","r_bench_binq1"):
    d=json.loads([l for l in open(f"gpurun_out/{f}.json") if l.startswith("{")][0]); print(f, d["ms_per_step"], d["roofline_knn"]["avg_launch_ms"])
PY
