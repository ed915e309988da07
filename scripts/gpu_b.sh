#!/bin/bash
# ncu --set full of the four TMA NTT kernels at shape C (and B)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
FHE_B200_NTT=tma timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tma -s 12 -c 4 -f -o gpurun_out/r2_tma_c python profiles/ntt_bench.py --shape C --reps 1 > gpurun_out/b_ncu_c.log 2>&1
echo "rc=$?" >> gpurun_out/b_ncu_c.log
FHE_B200_NTT=tma timeout 900 ncu --set full --clock-control none -k regex:ntt_tma -s 12 -c 4 -f -o gpurun_out/r2_tma_b python profiles/ntt_bench.py --shape B --reps 1 > gpurun_out/b_ncu_b.log 2>&1
echo "rc=$?" >> gpurun_out/b_ncu_b.log
tail -3 gpurun_out/b_ncu_c.log gpurun_out/b_ncu_b.log
