#!/bin/bash
mkdir -p gpurun_out
: > gpurun_out/ac_chunk.txt
for c in 64 128 256 512; do FHE_B200_CHUNK=$c timeout 300 python profiles/quick_bench.py 1024 3 >> gpurun_out/ac_chunk.txt 2>&1; done
cut -c1-200 gpurun_out/ac_chunk.txt
