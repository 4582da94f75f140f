#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 300 python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -3 | tee gpurun_out/quick1024.log
timeout 900 python bench.py --steps 5 2>&1 | tail -1 > gpurun_out/bench_ours.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_ours.json'))
print('value', d['value'], 'e2e', d['e2e']['value'], 'roofline', d['roofline']['frac'], 'chain', d['chain_roofline']['fp32']['frac'], d['kernel_share_of_step'])
PY
