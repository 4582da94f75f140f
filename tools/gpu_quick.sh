#!/bin/bash
# quick GPU check: whole GPU parity suite + default bench line (tag = $1)
TAG=${1:-quick}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$TAG.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_$TAG.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
tail -15 gpurun_out/pytest_gpu_$TAG.log
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$TAG.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'])
print({k:round(v['avg_launch_ms'],3) for k,v in d.get('kernels',{}).items()})
PY
