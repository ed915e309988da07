#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/p_e2e.txt
for cfg in "256 3" "512 3" "1024 3" "512 4"; do
  set -- $cfg
  FHE_BENCH_E2E_BATCH=$1 FHE_BENCH_E2E_SLOTS=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/p_tmp.json 2>gpurun_out/p_tmp.err
  python -c "
import json
d=json.loads(open('gpurun_out/p_tmp.json').read())
print('$cfg', 'e2e', round(d['e2e']['value'],1), d['e2e']['step_ms'], 'copy-only', d['e2e']['link_bound_products_per_s'], 'value', round(d['value'],1))" >> gpurun_out/p_e2e.txt 2>&1
done
cat gpurun_out/p_e2e.txt
