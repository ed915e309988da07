#!/bin/bash
# compute-sanitizer over the final kernels (pair rows kernel, fused tensor + inverse rows, cols, scaler, key-switch MAC, substitute, pack/unpack)
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 FHE_B200_NTT=tma
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_probe.py > gpurun_out/aj_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -3 gpurun_out/aj_memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python profiles/sanitize_probe.py > gpurun_out/aj_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/aj_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python profiles/sanitize_probe.py > gpurun_out/aj_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/aj_synccheck.txt
