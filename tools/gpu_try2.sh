#!/bin/bash
mkdir -p gpurun_out
for NB in 3 2 4; do
PMR446_FF_BLOCKS=$NB python bench.py --steps 10 --warmup 3 > gpurun_out/bench_ffnb$NB.json 2> gpurun_out/bench_ffnb$NB.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_ffnb$NB.json'))
print('NB=$NB ms_per_step', d['ms_per_step'], {k:round(v['avg_launch_ms'],3) for k,v in d.get('kernels',{}).items()})
PY
done
python -m pytest tests/test_gpu_pmr_parity.py -m gpu -x -q -k "cfg_b or awkward or large_chunk" 2>&1 | tail -2
