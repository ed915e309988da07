#!/bin/bash
# first GPU call of round 2: TMA NTT correctness, NTT timing per family, full GPU tests, short bench
mkdir -p gpurun_out
cd "$GRAFT_REPO_ROOT" 2>/dev/null || true
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt 2>&1
echo "== ntt tests under tma" > gpurun_out/a_log.txt
FHE_B200_NTT=tma timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_ntt_forward_backward" >> gpurun_out/a_log.txt 2>&1
echo "rc=$?" >> gpurun_out/a_log.txt
for fam in fast tma; do for shape in B C; do
  FHE_B200_NTT=$fam timeout 300 python profiles/ntt_bench.py --shape $shape >> gpurun_out/a_nttbench.txt 2>&1
done; done
echo "== full gpu tests" >> gpurun_out/a_log.txt
timeout 2400 python -m pytest tests -m gpu -x -q >> gpurun_out/a_log.txt 2>&1
echo "rc=$?" >> gpurun_out/a_log.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?" >> gpurun_out/a_log.txt
tail -5 gpurun_out/a_log.txt; cat gpurun_out/a_nttbench.txt; tail -c 1500 gpurun_out/a_bench.err
