#!/bin/bash
# chunks alternated over two side streams: parity (chunk-boundary test + set C) and A/B at the bench batch
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunk_boundary or full_size_set_c or alternate_code_paths" > gpurun_out/ak_tests.txt 2>&1; tail -2 gpurun_out/ak_tests.txt
{
FHE_B200_STREAMS=1 timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
FHE_B200_CHUNK=512 timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
FHE_B200_CHUNK=128 timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
FHE_B200_STREAMS=1 timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
timeout 300 python profiles/quick_bench.py 1024 3 2>&1 | tail -1
} > gpurun_out/ak_ab.txt
cat gpurun_out/ak_ab.txt
