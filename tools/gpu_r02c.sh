#!/bin/bash
# fused front end: bench + ncu full capture of the fused kernel (256 streams keeps the replay short)
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02c.json'))
print('ms_per_step', d['ms_per_step'], 'value', d['value'])
print({k:round(v['avg_launch_ms'],3) for k,v in d.get('kernels',{}).items()})
PY
python -m pytest tests/test_gpu_pmr_parity.py -m gpu -x -q -k "cfg_b or awkward or large_chunk" 2>&1 | tail -2
bash tools/gpu_prof6.sh fused_frontend ff_r02c python tools/quick_bench.py 256
