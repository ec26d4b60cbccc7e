#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lgssm or predict or sde or chain or smooth" > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2q_status.txt
timeout 300 python tools/prof_smooth_shared_one.py > gpurun_out/r2q_smooth.txt 2>&1
cat gpurun_out/r2q_status.txt; tail -5 gpurun_out/r2q_tests.log; cat gpurun_out/r2q_smooth.txt
