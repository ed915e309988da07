#!/bin/bash
# final verification of HEAD: GPU suite, smoke, default bench, reference arm
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/ai_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/ai_tests.txt
tail -4 gpurun_out/ai_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ai_smoke.txt 2>&1; tail -1 gpurun_out/ai_smoke.txt
T0=$(date +%s)
timeout 900 python bench.py > gpurun_out/ai_bench.json 2> gpurun_out/ai_bench.err; echo "bench rc=$? wall=$(( $(date +%s) - T0 ))s"
python -c "
import json
d=json.loads(open('gpurun_out/ai_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'e2e',d['e2e']['value'],'rot',d['secondary']['rotate']['value'],'frac',d['roofline']['frac'],'launches',d['gpu_launches'], d['verified']['bit_exact'], d['cpu_baseline']['value'], d['clocks'])"
