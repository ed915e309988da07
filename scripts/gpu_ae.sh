#!/bin/bash
# chunk-size sweep at the bench batch (1024 pairs), device-resident
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for c in 256 512 1024 128 256; do
  FHE_B200_CHUNK=$c timeout 600 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
done > gpurun_out/ae_chunk.txt
cat gpurun_out/ae_chunk.txt
