#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2c_status.txt
timeout 300 python tools/prof_scaled_grad.py > gpurun_out/r2c_scaled_grad.txt 2>&1; echo "prof rc=$?" >> gpurun_out/r2c_status.txt
timeout 300 python tools/latency_small.py > gpurun_out/r2c_latency_small.txt 2>&1; echo "lat rc=$?" >> gpurun_out/r2c_status.txt
timeout 300 python tools/prof_scaled_cfg5.py 3 > gpurun_out/r2c_cfg5.txt 2>&1; echo "cfg5 rc=$?" >> gpurun_out/r2c_status.txt
cat gpurun_out/r2c_status.txt; tail -30 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_scaled_grad.txt gpurun_out/r2c_latency_small.txt gpurun_out/r2c_cfg5.txt
