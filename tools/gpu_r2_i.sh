#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kalman_onepass_sweep.py > gpurun_out/r2i_sweep.txt 2>&1; echo "sweep rc=$?" > gpurun_out/r2i_status.txt
cat gpurun_out/r2i_status.txt; cat gpurun_out/r2i_sweep.txt
