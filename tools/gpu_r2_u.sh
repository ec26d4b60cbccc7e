#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched_restarts or chain or reference_example" > gpurun_out/r2u_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2u_status.txt
cat gpurun_out/r2u_status.txt; tail -12 gpurun_out/r2u_tests.log
