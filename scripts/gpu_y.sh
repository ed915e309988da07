#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mul_relin_against or full_size_set_c or set_b or set_c_across or golden or mixed_sizes" > gpurun_out/y_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/y_tests.txt
tail -3 gpurun_out/y_tests.txt
: > gpurun_out/y_quick.txt
for v in 22 13 14 22 14; do FHE_B200_TENSOR_V=$v timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/y_quick.txt 2>&1; done
FHE_B200_NO_TENSOR_FUSION=1 timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/y_quick.txt 2>&1
cut -c1-200 gpurun_out/y_quick.txt
