#!/bin/bash
# Per-kernel executed FP64 flops, DRAM bytes, pipe utilisation ... of every BASELINE config: one profiled call
# each (after the same command has exited 0 without ncu).  Output: gpurun_out/r2_costs_<case>.csv
# -> tools/ncu_costs.py turns them into profiles/r2_kernel_costs.json and profiles/r2_kernel_table.md
set -u
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,launch__registers_per_thread,sm__warps_active.avg.per_cycle_active
mkdir -p gpurun_out
for spec in "cfg1 1000000" "cfg2 33554432" "cfg2_ones 33554432" "cfg3_devroye 33554432" "cfg3_alternate 33554432" "cfg3_saddle 33554432" "cfg3_gamma 33554432" "cfg4 33554432" "pdf 33554432" "logpdf 33554432" "cdf 33554432" "logcdf 33554432"; do
  set -- $spec
  python tools/profile_case.py $1 $2 1 > gpurun_out/plain_costs_$1.log 2>&1 || { echo "$1 failed without ncu"; continue; }
  ncu --metrics $M --clock-control none --kernel-name-base demangled --csv --log-file gpurun_out/r2_costs_$1.csv \
      python tools/profile_case.py $1 $2 1 > gpurun_out/ncu_costs_$1.log 2>&1
  echo "$1: $(grep -c gpu__time_duration gpurun_out/r2_costs_$1.csv) kernel rows"
done
