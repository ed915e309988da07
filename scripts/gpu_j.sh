#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mul_relin_against or full_size_set_c or set_b or set_c_across or golden or mixed_modulus or mul_general" > gpurun_out/j_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/j_tests.txt
tail -4 gpurun_out/j_tests.txt
: > gpurun_out/j_quick.txt
timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/j_quick.txt 2>&1
FHE_B200_NO_TENSOR_FUSION=1 timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/j_quick.txt 2>&1
cat gpurun_out/j_quick.txt
