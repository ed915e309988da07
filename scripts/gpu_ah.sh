#!/bin/bash
mkdir -p gpurun_out
cd bench_micro && timeout 300 ./mac_dfma > ../gpurun_out/ah_mac_dfma.txt 2>&1; cd ..
cat gpurun_out/ah_mac_dfma.txt
