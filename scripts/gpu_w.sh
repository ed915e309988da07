#!/bin/bash
# final verification + final profile artifacts
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/w_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/w_tests.txt
tail -3 gpurun_out/w_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/w_smoke.txt 2>&1; tail -1 gpurun_out/w_smoke.txt
timeout 900 python bench.py > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/w_launches.csv python profiles/probe.py mulrelin 64 > gpurun_out/w_probe.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tma -s 8 -c 4 -f -o gpurun_out/r2_final2_ntt14 python profiles/probe.py ntt14 > gpurun_out/w_ntt14.log 2>&1
python -c "
import json
d=json.loads(open('gpurun_out/w_bench.json').read())
print('value',d['value'],'e2e',d['e2e']['value'],d['e2e']['batch'],'rot',d['secondary']['rotate']['value'],'frac',d['roofline']['frac'], d['roofline']['issue_roofline']['frac'], 'launches', d['gpu_launches'])
print(d['vs_single_thread'], d['clocks'], d['e2e']['link_bound_products_per_s'])"
python profiles/launch_summary.py gpurun_out/w_launches.csv 2>/dev/null | head -14
