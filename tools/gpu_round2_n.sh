#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "unordered or knn" 2>&1 | tail -4 > gpurun_out/n_pytest.txt
: > gpurun_out/n_sweep.jsonl
for tau in 0.45 0.6 0.8 1.0; do
  echo "# P2W_SEL_TAU=$tau" >> gpurun_out/n_sweep.jsonl
  P2W_SEL_TAU=$tau VOTE_CELLS=0.05,0.07,0.09,0.12 timeout 300 python tools/bench_knn.py vote 2>&1 | cut -c1-700 >> gpurun_out/n_sweep.jsonl
done
tail -3 gpurun_out/n_pytest.txt; python - <<'PY'
import json
for l in open("gpurun_out/n_sweep.jsonl"):
    if l.startswith("#"): print(l.strip()); continue
    try: d=json.loads(l)
    except Exception: print(l[:200]); continue
    print(d["case"][:60], d["ms"], d.get("evals_per_query"), d.get("select"))
PY
