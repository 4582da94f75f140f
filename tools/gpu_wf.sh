#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "waterfall or shim or golden or harness" 2>&1 | tail -3
echo "fast:"; python tools/probe_waterfall.py
echo "generic:"; PMR446_WATERFALL=generic python tools/probe_waterfall.py
echo "W=128 fast:"; WF_W=128 python tools/probe_waterfall.py
