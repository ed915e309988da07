#!/bin/bash
# occupancy sweep of the TMA NTT kernels (ring depth x CTAs per SM)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/c_sweep.txt
for rows in 44 34 35 26; do for cols in 3 2; do
  echo "== rows $rows cols $cols" >> gpurun_out/c_sweep.txt
  FHE_B200_NTT=tma FHE_B200_TMA_ROWS=$rows FHE_B200_TMA_COLS=$cols timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_ntt_forward_backward and (13 or 14 or 15)" 2>&1 | tail -1 >> gpurun_out/c_sweep.txt
  for shape in B C; do
    FHE_B200_NTT=tma FHE_B200_TMA_ROWS=$rows FHE_B200_TMA_COLS=$cols timeout 300 python profiles/ntt_bench.py --shape $shape >> gpurun_out/c_sweep.txt 2>&1
  done
done; done
cat gpurun_out/c_sweep.txt | cut -c1-330
