#!/bin/bash
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active,launch__registers_per_thread
timeout 900 ncu --metrics $M --clock-control none -c 900 --csv --log-file gpurun_out/r2o_launches_scaled_grad.csv python tools/prof_scaled_grad.py > gpurun_out/r2o_ncu.log 2>&1; echo "ncu rc=$?" > gpurun_out/r2o_status.txt
cat gpurun_out/r2o_status.txt; tail -4 gpurun_out/r2o_ncu.log
