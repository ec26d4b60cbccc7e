#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lgssm" > gpurun_out/r2l_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2l_status.txt
timeout 900 python tools/kalman_grad_time.py > gpurun_out/r2l_grad.txt 2>&1; echo "grad rc=$?" >> gpurun_out/r2l_status.txt
cat gpurun_out/r2l_status.txt; tail -15 gpurun_out/r2l_tests.log; cat gpurun_out/r2l_grad.txt
