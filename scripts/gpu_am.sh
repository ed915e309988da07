#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk_runner or chunk_boundary or mul_general or custom_multiplication or galois_and_key_switch or golden" > gpurun_out/am_tests.txt 2>&1; tail -12 gpurun_out/am_tests.txt
