#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -q > gpurun_out/r2g_2gpu_group_tests.log 2>&1; echo "group tests rc=$?" > gpurun_out/r2g_2gpu_status.txt
cat gpurun_out/r2g_2gpu_status.txt; tail -n 6 gpurun_out/r2g_2gpu_group_tests.log | cut -c1-300
