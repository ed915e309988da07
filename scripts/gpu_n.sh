#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_devices" > gpurun_out/n_twodev.txt 2>&1; tail -2 gpurun_out/n_twodev.txt
for wc in 0 1; do
FHE_BENCH_WC=$wc timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2961$wc bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/n_bench8_wc$wc.json 2> gpurun_out/n_bench8_wc$wc.err; echo "bench8 wc=$wc rc=$?"
python -c "
import json,sys
d=json.loads(open('gpurun_out/n_bench8_wc$wc.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e']['step_ms'], 'agg', d['e2e'].get('aggregate_h2d_gbs'), 'copy-only products/s', d['e2e'].get('link_bound_products_per_s'))
print(d['e2e']['concurrent_pinned_copy_gbs_per_rank'][0], d['e2e']['concurrent_pinned_copy_gbs_per_rank'][7])"
done
