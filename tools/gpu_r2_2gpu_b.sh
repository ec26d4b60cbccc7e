#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2_2gpu_group_tests.log 2>&1; echo "group tests rc=$?" > gpurun_out/r2_2gpu_status.txt
cat gpurun_out/r2_2gpu_status.txt; tail -4 gpurun_out/r2_2gpu_group_tests.log
