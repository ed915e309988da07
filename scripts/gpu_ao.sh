#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_cpp_host.py -m gpu -x -q > gpurun_out/ao_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/ao_tests.txt; tail -3 gpurun_out/ao_tests.txt
