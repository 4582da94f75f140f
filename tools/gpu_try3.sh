#!/bin/bash
for cfg in "PMR446_FF_THREADS=128" "PMR446_FF_THREADS=64" "PMR446_FF_THREADS=32" "PMR446_FF_SEG=3072" "PMR446_FF_SEG=12288" "PMR446_FF_SEG=12288 PMR446_FF_THREADS=64"; do
  env $cfg python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -1 | sed "s/^/$cfg /"
done
