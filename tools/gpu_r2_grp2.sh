#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -q -k "group or sharded or fit_through or slice" > gpurun_out/r2grp_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2grp_status.txt
cat gpurun_out/r2grp_status.txt; tail -25 gpurun_out/r2grp_tests.log | cut -c1-400
