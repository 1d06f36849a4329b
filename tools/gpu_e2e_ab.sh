#!/bin/bash
for v in 1 0 1 0; do
P2W_DENSE_TC=$v timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('dense_tc=$v', round(d['ms_per_step'],2), 'e2e ms', round(1e9/d['e2e']['value'],2))"
done
nvidia-smi --query-gpu=pcie.link.gen.current,pcie.link.width.current --format=csv
