#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/q_e2e.txt
for t in 1 0; do
  FHE_BENCH_E2E_TAPER=$t timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/q_tmp.json 2>gpurun_out/q_tmp.err
  python -c "
import json
d=json.loads(open('gpurun_out/q_tmp.json').read())
print('taper $t', 'e2e', round(d['e2e']['value'],1), d['e2e']['step_ms'], d['e2e']['chunks'], 'copy-only', d['e2e']['link_bound_products_per_s'], 'value', round(d['value'],1))" >> gpurun_out/q_e2e.txt 2>&1
  tail -c 300 gpurun_out/q_tmp.err >> gpurun_out/q_e2e.txt
done
cat gpurun_out/q_e2e.txt
