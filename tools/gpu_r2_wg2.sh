#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_parity.py -m gpu -q -s -k "ill_conditioned or scaled or sharded or group" > gpurun_out/r2wg2_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2wg2_status.txt
cat gpurun_out/r2wg2_status.txt; tail -25 gpurun_out/r2wg2_tests.log | cut -c1-300
