#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/l_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/l_tests.txt
tail -3 gpurun_out/l_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.txt 2>&1; tail -1 gpurun_out/l_smoke.txt
timeout 900 python bench.py > gpurun_out/l_bench.json 2> gpurun_out/l_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/l_bench_ref.json 2> gpurun_out/l_bench_ref.err; echo "ref rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/l_bench.json').read())
print('value',d['value'],'e2e',d['e2e']['value'],'rot',d['secondary']['rotate']['value'],'frac',d['roofline']['frac'], d['roofline']['issue_roofline']['frac'], 'launches', d['gpu_launches'])
print(d['vs_single_thread'], d['clocks'])
r=json.loads(open('gpurun_out/l_bench_ref.json').read()); print('ref', r['value'], r['cpu_baseline']['cores'])"
tail -c 300 gpurun_out/l_bench.err
