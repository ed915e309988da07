#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "galois or mul_relin_against or full_size_set_c or set_b or set_c_across or leveled or rgsw or golden" > gpurun_out/g_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.txt
tail -4 gpurun_out/g_tests.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/g_launches.csv python profiles/probe.py mulrelin 64 > gpurun_out/g_probe.log 2>&1
python profiles/launch_summary.py gpurun_out/g_launches.csv 2>/dev/null | head -12
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
python -c "
import json
d=json.loads(open('gpurun_out/g_bench.json').read())
print('value',d['value'],'e2e',d['e2e']['value'],'rot',d['secondary']['rotate']['value'],'frac',d['roofline']['frac'])
print(d['e2e'].get('concurrent_pinned_copy_gbs_per_rank'), d.get('result_gather'))"
tail -c 400 gpurun_out/g_bench.err
