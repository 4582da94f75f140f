#!/bin/bash
T=${1:-cfgs}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_$T.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu_$T.log; tail -3 gpurun_out/pytest_gpu_$T.log
for c in cfg1 cfg2 receiver; do python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_${T}_$c.json 2> gpurun_out/bench_${T}_$c.err; python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${T}_$c.json')); print('$c', round(d['ms_per_step'],3), 'ms', round(d['value']), 'Msps')
except Exception as e:
    print('$c FAILED', e); print(open('gpurun_out/bench_${T}_$c.err').read()[-600:])
PY
done
