#!/bin/bash
# tools/quick_bench.py (1024 streams, per-kernel timings) once per environment setting ("-" = none): gpu_quick_env.sh tag "VAR=a" "VAR=b OTHER=c" ...
TAG=$1; shift
mkdir -p gpurun_out
i=0
for e in "$@"; do
  [ "$e" = "-" ] && e=""
  env $e python tools/quick_bench.py --streams 1024 --steps 6 > gpurun_out/quick_${TAG}_$i.log 2>&1
  echo "[$e] $(grep 'step 7' gpurun_out/quick_${TAG}_$i.log | cut -d' ' -f2-4) $(tail -1 gpurun_out/quick_${TAG}_$i.log)"
  i=$((i+1))
done
