#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/k_quick.txt
for v in 22 13 14; do FHE_B200_TENSOR_V=$v timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/k_quick.txt 2>&1; done
cat gpurun_out/k_quick.txt
