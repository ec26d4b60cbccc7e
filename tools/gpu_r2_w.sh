#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2w_status.txt
cat gpurun_out/r2w_status.txt; tail -6 gpurun_out/r2w_tests.log
