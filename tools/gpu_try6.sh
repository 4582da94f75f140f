#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python tools/quick_bench.py --streams 1024 --fs 1024000 --chunk 1000000 --steps 3 2>&1 | tail -1
PMR446_FRONTEND=split python tools/quick_bench.py --streams 1024 --fs 1024000 --chunk 1000000 --steps 3 2>&1 | tail -1 | sed 's/^/split: /'
python tools/quick_bench.py --streams 1024 --steps 3 2>&1 | tail -1
