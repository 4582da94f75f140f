#!/bin/bash
# usage: gpu_prof6.sh <kernel-regex> <tag> <python script + args>
mkdir -p gpurun_out
K=$1; T=$2; shift 2
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$K" --launch-skip 2 --launch-count 1 -o gpurun_out/prof_$T -f "$@" > gpurun_out/ncu_full_$T.log 2>&1
tail -2 gpurun_out/ncu_full_$T.log
