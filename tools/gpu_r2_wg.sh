#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "ill_conditioned or scaled_dtc_grad or dtc_grad or golden" > gpurun_out/r2wg_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2wg_status.txt
timeout 300 python tools/prof_whitened_grad.py > gpurun_out/r2wg_time.txt 2>&1
cat gpurun_out/r2wg_status.txt; tail -15 gpurun_out/r2wg_tests.log; cat gpurun_out/r2wg_time.txt | cut -c1-150
