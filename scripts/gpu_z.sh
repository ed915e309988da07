#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mul_relin_against or full_size_set_c or set_b or set_c_across or golden or mixed_sizes or custom_mult" > gpurun_out/z_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/z_tests.txt
tail -3 gpurun_out/z_tests.txt
: > gpurun_out/z_quick.txt
for rep in 1 2; do timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/z_quick.txt 2>&1; done
cut -c1-200 gpurun_out/z_quick.txt
