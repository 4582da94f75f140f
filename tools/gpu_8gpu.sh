#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_r02x_${N}gpu.json 2> gpurun_out/bench_r02x_${N}gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02x_${N}gpu.json'))
print('${N}gpu value', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']), d['e2e'].get('pcie_ceiling'))
print([ (round(r['h2d_gbs_all_ranks_copying'],1), round(r['e2e_input_gbs'],1), r['numa_node']) for r in d['per_rank']])
PY
