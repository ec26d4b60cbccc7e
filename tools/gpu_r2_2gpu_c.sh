#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2c_2gpu_status.txt
timeout 600 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2c_2gpu_group_tests.log 2>&1; echo "group tests rc=$?" >> gpurun_out/r2c_2gpu_status.txt
timeout 300 python tools/bench_sharded_scaled.py --devices 2 --npoints 1000000 --pseudo 1024 --dim 1 > gpurun_out/r2c_sharded_scaled_2gpu_1m.json 2> gpurun_out/r2c_sharded.err; echo "sharded 1M rc=$?" >> gpurun_out/r2c_2gpu_status.txt
timeout 400 python tools/bench_sharded_scaled.py --devices 2 > gpurun_out/r2c_sharded_scaled_2gpu.json 2>> gpurun_out/r2c_sharded.err; echo "sharded cfg5 rc=$?" >> gpurun_out/r2c_2gpu_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --fit-iterations 1 > gpurun_out/r2c_2gpu_bench.json 2> gpurun_out/r2c_2gpu_bench.err; echo "bench rc=$?" >> gpurun_out/r2c_2gpu_status.txt
cat gpurun_out/r2c_2gpu_status.txt; tail -4 gpurun_out/r2c_2gpu_group_tests.log; cat gpurun_out/r2c_sharded_scaled_2gpu_1m.json gpurun_out/r2c_sharded_scaled_2gpu.json; tail -q -n 3 gpurun_out/r2c_sharded.err gpurun_out/r2c_2gpu_bench.err
