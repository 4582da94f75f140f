#!/bin/bash
# usage: gpu_ncu_set.sh <tag> <A|B>   -- ncu full captures, split in two calls to stay under gpurun's 64 MiB return limit
T=$1
mkdir -p gpurun_out
if [ "$2" = "A" ]; then
bash tools/gpu_prof6.sh fused_frontend ff_$T python tools/quick_bench.py --streams 512 --steps 1
bash tools/gpu_prof6.sh channelize16 ch_$T python tools/quick_bench.py --streams 512 --steps 1
bash tools/gpu_prof6.sh audio_fft af_$T python tools/quick_bench.py --streams 512 --steps 1
else
WF_STREAMS=256 bash tools/gpu_prof6.sh wf_accumulate_fast wf_$T python tools/probe_waterfall.py
DSD_STREAMS=256 bash tools/gpu_prof6.sh front6 f6_$T python tools/probe_dsd.py
bash tools/gpu_prof6.sh channelize_generic_tile cg_$T python tools/probe_wideband.py
bash tools/gpu_prof6.sh wf_accumulate_kernel wfb_$T python tools/probe_wideband.py
fi
ls -la gpurun_out/*.ncu-rep
