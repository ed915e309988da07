#!/bin/bash
# 8-GPU box: two-device test in one process, then the bench at N = 8 (own arm)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi topo -m > gpurun_out/m_topo.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices" > gpurun_out/m_twodev.txt 2>&1; tail -2 gpurun_out/m_twodev.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/m_bench8.json 2> gpurun_out/m_bench8.err; echo "bench8 rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/m_bench8.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], 'agg', d['e2e'].get('aggregate_h2d_gbs'), d['e2e'].get('link_bound_products_per_s'))
for r in d['e2e']['concurrent_pinned_copy_gbs_per_rank']: print(r)
print(d['result_gather'])"
tail -c 500 gpurun_out/m_bench8.err
