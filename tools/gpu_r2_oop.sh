#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_group.py -m gpu -q -k "ill_conditioned or row_sharded_loopback" > gpurun_out/r2oop_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2oop_status.txt
timeout 200 python tools/prof_whitened_grad.py > gpurun_out/r2oop_time.txt 2>&1
cat gpurun_out/r2oop_status.txt; tail -n 4 gpurun_out/r2oop_tests.log | cut -c1-300; grep -E "value|plain" gpurun_out/r2oop_time.txt | cut -c1-160
