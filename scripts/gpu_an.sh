#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 140 python -m pytest tests -m gpu -x -q -k "not alternate_code_paths and not full_size and not chunk_boundary and not chunk_runner and not set_b and not mul_general and not custom_multiplication and not galois_and_key_switch and not golden" > gpurun_out/an_tests.txt 2>&1; echo "rc=$?" >> gpurun_out/an_tests.txt; tail -5 gpurun_out/an_tests.txt
