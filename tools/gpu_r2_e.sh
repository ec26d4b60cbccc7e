#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2e_status.txt
timeout 300 python tools/latency_small.py > gpurun_out/r2e_latency_small.txt 2>&1; echo "lat rc=$?" >> gpurun_out/r2e_status.txt
cat gpurun_out/r2e_status.txt; tail -30 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_latency_small.txt
