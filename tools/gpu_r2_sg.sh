#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scaled" > gpurun_out/r2sg_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2sg_status.txt
timeout 300 python tools/prof_scaled_grad.py > gpurun_out/r2sg_grad.txt 2>&1
cat gpurun_out/r2sg_status.txt; tail -3 gpurun_out/r2sg_tests.log; grep "value+grad" gpurun_out/r2sg_grad.txt | cut -c1-80
