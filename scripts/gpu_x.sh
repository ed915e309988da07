#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:tensor_rows" -c 1 -f -o gpurun_out/r2_tensor_rows python profiles/probe.py mulrelin 32 > gpurun_out/x.log 2>&1
tail -2 gpurun_out/x.log
