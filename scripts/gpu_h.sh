#!/bin/bash
# ncu --set full captures of the shipped kernels (for profiles/): bench-shape NTT (N=2^14, [256][8]) and one mul+relin
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ntt_tma -s 8 -c 4 -f -o gpurun_out/r2_final_ntt14 python profiles/probe.py ntt14 > gpurun_out/h_ntt14.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:scale_tma|ksmac_tma|tensor_kernel" -c 7 -f -o gpurun_out/r2_final_mulrelin python profiles/probe.py mulrelin 32 > gpurun_out/h_mulrelin.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2_launches_rotate_64ct.csv python profiles/probe.py rotate 64 > gpurun_out/h_rot.log 2>&1
tail -2 gpurun_out/h_ntt14.log gpurun_out/h_mulrelin.log gpurun_out/h_rot.log
