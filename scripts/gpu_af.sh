#!/bin/bash
# final build at N=2: the multi-rank bench path (torchrun, NCCL counters + checksum gather)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/af_bench2.json 2> gpurun_out/af_bench2.err; echo "bench2 rc=$? wall=$(( $(date +%s) - T0 ))s"
python -c "
import json
d=json.loads(open('gpurun_out/af_bench2.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['batch'], 'copy-only products/s', d['e2e'].get('link_bound_products_per_s'), 'rot', d['secondary']['rotate']['value'])
print(d['verified'], d['result_gather']['collective'], d['clocks'])"
tail -c 300 gpurun_out/af_bench2.err
