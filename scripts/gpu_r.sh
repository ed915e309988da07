#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mixed_sizes_through or mixed_modulus" > gpurun_out/r_tests.txt 2>&1; tail -3 gpurun_out/r_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
: > gpurun_out/r_roll.txt
for roll in 0 1; do
  FHE_B200_COLS_ROLL=$roll FHE_B200_NTT=tma timeout 300 python profiles/ntt_bench.py --shape C >> gpurun_out/r_roll.txt 2>&1
  FHE_B200_COLS_ROLL=$roll timeout 300 python profiles/quick_bench.py 256 4 >> gpurun_out/r_roll.txt 2>&1
done
cut -c1-260 gpurun_out/r_roll.txt
