#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
{
for sN in 2 3 4; do FHE_B200_STREAMS=$sN timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1; done
FHE_B200_STREAMS=4 FHE_B200_CHUNK=512 timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
FHE_B200_STREAMS=3 FHE_B200_CHUNK=384 timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
} > gpurun_out/al_ab.txt
cat gpurun_out/al_ab.txt
