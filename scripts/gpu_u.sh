#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "galois or full_size_set_c or set_b or set_c_across or mixed_sizes or mul_relin_against" > gpurun_out/u_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/u_tests.txt
tail -3 gpurun_out/u_tests.txt
: > gpurun_out/u_quick.txt
for rep in 1 2; do
timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/u_quick.txt 2>&1
FHE_B200_NO_DIAG_SKIP=1 timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/u_quick.txt 2>&1
FHE_B200_LIB=$PWD/fhe_rs_b200/libfhe_b200_prev.so timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/u_quick.txt 2>&1
done
cat gpurun_out/u_quick.txt
