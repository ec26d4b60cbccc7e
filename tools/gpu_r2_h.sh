#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:kf_chunk_element -s 1 -c 1 -o gpurun_out/r2h_kf1 -f python tools/prof_kalman_1x10m.py > gpurun_out/r2h_ncu.log 2>&1; echo "ncu rc=$?" > gpurun_out/r2h_status.txt
cat gpurun_out/r2h_status.txt; tail -5 gpurun_out/r2h_ncu.log
