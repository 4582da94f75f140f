#!/bin/bash
# final round-2 evidence on the shipped kernels: suite, bench lines of every config, launch list, fused front end ncu report, soak past 2^32 samples
T=${1:-r02x}
bash tools/gpu_r02w.sh $T
bash tools/gpu_prof6.sh fused_frontend ff_$T python tools/quick_bench.py --streams 512 --steps 1
SOAK_CHUNKS=1850 python tools/soak_run.py > gpurun_out/soak_$T.log 2>&1; tail -2 gpurun_out/soak_$T.log
