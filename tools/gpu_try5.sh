#!/bin/bash
for cfg in "PMR446_CH_PIPE=0" "PMR446_CH_PIPE=1"; do
  env $cfg python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -1 | sed "s/^/$cfg /"
done
python -m pytest tests/test_gpu_pmr_parity.py tests/test_golden.py -m gpu -x -q 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_rx_r02.csv python tools/probe_receiver.py | tail -1
