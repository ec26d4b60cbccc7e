#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2d_launches_value.csv python tools/prof_scaled_value_one.py > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?" > gpurun_out/r2d_status.txt
timeout 300 python tools/latency_small.py > gpurun_out/r2d_latency_small.txt 2>&1; echo "lat rc=$?" >> gpurun_out/r2d_status.txt
cat gpurun_out/r2d_status.txt gpurun_out/r2d_latency_small.txt
