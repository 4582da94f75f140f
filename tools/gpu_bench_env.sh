#!/bin/bash
# bench.py once per environment setting ("-" = none): tools/gpu_bench_env.sh tag "VAR=a" "VAR=b OTHER=c" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for e in "$@"; do
  [ "$e" = "-" ] && e=""
  env $e python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_$i.json 2> gpurun_out/bench_${TAG}_$i.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${TAG}_$i.json'))
print('[$e]', 'ms_per_step', round(d['ms_per_step'],4), {k:round(v['avg_launch_ms'],3) for k,v in d.get('kernels',{}).items()})
PY
  i=$((i+1))
done
