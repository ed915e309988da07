#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1 FHE_B200_NTT=tma
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python profiles/sanitize_probe.py > gpurun_out/s_memcheck.txt 2>&1; echo "memcheck rc=$?"; tail -5 gpurun_out/s_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report analysis --print-limit 20 python profiles/sanitize_probe.py > gpurun_out/s_racecheck.txt 2>&1; echo "racecheck rc=$?"; tail -8 gpurun_out/s_racecheck.txt
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python profiles/sanitize_probe.py > gpurun_out/s_synccheck.txt 2>&1; echo "synccheck rc=$?"; tail -4 gpurun_out/s_synccheck.txt
