#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "scaled" > gpurun_out/r2n_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2n_status.txt
timeout 600 python tools/latency_small.py > gpurun_out/r2n_latency_small.txt 2>&1; echo "lat rc=$?" >> gpurun_out/r2n_status.txt
cat gpurun_out/r2n_status.txt; tail -15 gpurun_out/r2n_tests.log; cat gpurun_out/r2n_latency_small.txt
