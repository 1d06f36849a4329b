#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_vote.py tests/test_gpu_engine.py tests/test_gpu_fullsize.py tests/test_gpu_distributed.py tests/test_gpu_model.py -x -q 2>&1 | tail -12 > gpurun_out/m_pytest.txt
timeout 300 python tools/bench_knn.py vote > gpurun_out/m_knn_vote.jsonl 2>&1
P2W_KNN_SELECT=0 timeout 300 python tools/bench_knn.py vote >> gpurun_out/m_knn_vote.jsonl 2>&1
timeout 600 python tools/step_timeline.py 1000000 14 > gpurun_out/m_timeline.txt 2>&1
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench.json 2>/dev/null
P2W_KNN_SELECT=0 timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/m_bench_noselect.json 2>/dev/null
tail -6 gpurun_out/m_pytest.txt; cut -c1-330 gpurun_out/m_knn_vote.jsonl; head -20 gpurun_out/m_timeline.txt | cut -c1-120; cut -c1-200 gpurun_out/m_bench.json; cut -c1-200 gpurun_out/m_bench_noselect.json
