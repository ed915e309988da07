#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices or ntt_forward_backward" > gpurun_out/ab_tests.txt 2>&1; tail -2 gpurun_out/ab_tests.txt
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/ab_bench8.json 2> gpurun_out/ab_bench8.err; echo "bench8 rc=$? wall=$(( $(date +%s) - T0 ))s"
python -c "
import json
d=json.loads(open('gpurun_out/ab_bench8.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['batch'], d['e2e']['step_ms'], 'copy-only products/s', d['e2e'].get('link_bound_products_per_s'), 'rot', d['secondary']['rotate']['value'])
print(d['verified'], d['result_gather']['collective'])"
T0=$(date +%s)
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29634 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/ab_ref8.json 2> gpurun_out/ab_ref8.err; echo "ref8 rc=$? wall=$(( $(date +%s) - T0 ))s"; tail -c 300 gpurun_out/ab_ref8.json
tail -c 400 gpurun_out/ab_bench8.err
