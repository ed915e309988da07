#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scalers or mul_relin_against or full_size_set_c or set_b or custom_multiplication or golden or packed_mul or two_devices" > gpurun_out/f_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/f_tests.txt
tail -4 gpurun_out/f_tests.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches.csv python profiles/probe.py mulrelin 64 > gpurun_out/f_probe.log 2>&1
python profiles/launch_summary.py gpurun_out/f_launches.csv | head -12
FHE_B200_SCALER=classic timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f_launches_classic.csv python profiles/probe.py mulrelin 64 > gpurun_out/f_probe2.log 2>&1
python profiles/launch_summary.py gpurun_out/f_launches_classic.csv | head -4
