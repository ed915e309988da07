#!/bin/bash
# message layer (protobuf framing + device pack/unpack), power-basis substitute, C++ wire header: full GPU suite + smoke
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/ad_tests.txt 2>&1; echo "tests rc=$?" >> gpurun_out/ad_tests.txt
tail -25 gpurun_out/ad_tests.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/ad_smoke.txt 2>&1; tail -1 gpurun_out/ad_smoke.txt
