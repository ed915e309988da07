#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FHE_B200_NTT=tma FHE_B200_TMA_ROWS=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ntt_forward_backward or set_b or full_size_set_c or mixed_sizes" > gpurun_out/v_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/v_tests.txt
tail -3 gpurun_out/v_tests.txt
: > gpurun_out/v_quick.txt
for rep in 1 2; do for v in 44 2; do
  FHE_B200_TMA_ROWS=$v timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/v_quick.txt 2>&1
done; done
FHE_B200_TMA_ROWS=2 FHE_B200_NTT=tma timeout 300 python profiles/ntt_bench.py --shape C >> gpurun_out/v_quick.txt 2>&1
cut -c1-250 gpurun_out/v_quick.txt
