#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -4
python -m pytest tests/test_gpu_shard_nccl.py -m gpu -x -q > gpurun_out/pytest_nccl_2gpu_r02.log 2>&1; echo "exit $?" >> gpurun_out/pytest_nccl_2gpu_r02.log; tail -4 gpurun_out/pytest_nccl_2gpu_r02.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02_2gpu.json 2> gpurun_out/bench_r02_2gpu.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_2gpu.json'))
print('2gpu', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('pcie_ceiling'))
print([ (r['h2d_gbs_all_ranks_copying'], r['e2e_input_gbs'], r['pcie_gen'], r['pcie_width'], r['numa_node']) for r in d['per_rank']])
PY
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --scaling strong > gpurun_out/bench_r02_2gpu_strong.json 2> gpurun_out/bench_r02_2gpu_strong.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_r02_2gpu_strong.json'))
print('2gpu strong', d['value'], d['ms_per_step'], d['config'].get('streams_per_gpu'), 'e2e', d['e2e']['value'])
PY
