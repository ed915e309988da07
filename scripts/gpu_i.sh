#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
: > gpurun_out/i_sweep.txt
for ks in 2 3 4; do for un in 2 4; do
  FHE_B200_KS_STAGES=$ks FHE_B200_SCALE_UNROLL=$un timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/i_sweep.txt 2>&1
done; done
FHE_B200_KSMAC=classic FHE_B200_SCALER=classic FHE_B200_NTT=fast timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/i_sweep.txt 2>&1
cat gpurun_out/i_sweep.txt
