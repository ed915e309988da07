#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FHE_B200_NTT=tma FHE_B200_COLS_ROLL=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt_forward_backward or set_b or full_size_set_c" > gpurun_out/aa_tests.txt 2>&1; tail -2 gpurun_out/aa_tests.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt_forward_backward or set_b" > gpurun_out/aa_tests2.txt 2>&1; tail -2 gpurun_out/aa_tests2.txt
: > gpurun_out/aa_quick.txt
for rep in 1 2; do for roll in 0 1; do
  FHE_B200_COLS_ROLL=$roll timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/aa_quick.txt 2>&1
done; done
cut -c1-200 gpurun_out/aa_quick.txt
