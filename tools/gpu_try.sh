#!/bin/bash
# quick A/B: default bench line + three 2.4 Msps parity tests (tag = $1)
TAG=${1:-try}
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'])
print({k:round(v['avg_launch_ms'],3) for k,v in d.get('kernels',{}).items()})
PY
python -m pytest tests/test_gpu_pmr_parity.py -m gpu -x -q -k "cfg_b or awkward or large_chunk" 2>&1 | tail -2
