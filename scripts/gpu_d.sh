#!/bin/bash
mkdir -p gpurun_out
cd bench_micro
./bf_rf > ../gpurun_out/d_bf_rf.txt 2>&1
timeout 600 ncu --metrics sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,sm__cycles_elapsed.avg,smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio,smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio --clock-control none --csv --log-file ../gpurun_out/d_bf_rf_ncu.csv ./bf_rf > /dev/null 2>&1
cat ../gpurun_out/d_bf_rf.txt
