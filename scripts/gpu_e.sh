#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/e_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/e_tests.txt
: > gpurun_out/e_nttbench.txt
for shape in B C; do FHE_B200_NTT=tma timeout 300 python profiles/ntt_bench.py --shape $shape >> gpurun_out/e_nttbench.txt 2>&1; done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_mulrelin_64ct.csv python profiles/probe.py mulrelin 64 > gpurun_out/e_probe.log 2>&1
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
tail -3 gpurun_out/e_tests.txt; cat gpurun_out/e_nttbench.txt | cut -c1-300; tail -c 600 gpurun_out/e_bench.err
