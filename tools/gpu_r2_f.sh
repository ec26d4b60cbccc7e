#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lgssm" > gpurun_out/r2f_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2f_status.txt
timeout 900 python tools/kalman_onepass_sweep.py > gpurun_out/r2f_sweep.txt 2>&1; echo "sweep rc=$?" >> gpurun_out/r2f_status.txt
cat gpurun_out/r2f_status.txt; tail -15 gpurun_out/r2f_tests.log; cat gpurun_out/r2f_sweep.txt
