#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "galois or full_size_set_c or set_b or set_c_across or leveled or rgsw or golden or mixed_sizes or inner_sum or expand" > gpurun_out/t_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/t_tests.txt
tail -3 gpurun_out/t_tests.txt
: > gpurun_out/t_quick.txt
timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/t_quick.txt 2>&1
FHE_B200_NO_DIAG_SKIP=1 timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/t_quick.txt 2>&1
cat gpurun_out/t_quick.txt
