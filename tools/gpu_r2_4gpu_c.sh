#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_4gpu_status.txt
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2c_4gpu_group_tests.log 2>&1; echo "group tests rc=$?" >> gpurun_out/r2c_4gpu_status.txt
timeout 500 python tools/bench_sharded_scaled.py --devices 4 > gpurun_out/r2c_sharded_scaled_4gpu.json 2> gpurun_out/r2c_sharded4.err; echo "sharded cfg5 rc=$?" >> gpurun_out/r2c_4gpu_status.txt
cat gpurun_out/r2c_4gpu_status.txt; tail -4 gpurun_out/r2c_4gpu_group_tests.log; cat gpurun_out/r2c_sharded_scaled_4gpu.json; tail -n 3 gpurun_out/r2c_sharded4.err
