#!/bin/bash
# final build at N=8 (and N=4 on the same box): refresh of the multi-GPU records
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for n in 8 4; do
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2965$n bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/ag_bench$n.json 2> gpurun_out/ag_bench$n.err; echo "bench$n rc=$? wall=$(( $(date +%s) - T0 ))s"
python -c "
import json
d=json.loads(open('gpurun_out/ag_bench$n.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['batch'], 'copy-only products/s', d['e2e'].get('link_bound_products_per_s'), 'rot', d['secondary']['rotate']['value'])
print(d['verified'], d['clocks'])"
done
