#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r2_2gpu_status.txt
timeout 900 python -m pytest tests/test_gpu_group.py -m gpu -x -q > gpurun_out/r2_2gpu_group_tests.log 2>&1; echo "group tests rc=$?" >> gpurun_out/r2_2gpu_status.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_2gpu_bench.json 2> gpurun_out/r2_2gpu_bench.err; echo "bench rc=$?" >> gpurun_out/r2_2gpu_status.txt
cat gpurun_out/r2_2gpu_status.txt; tail -5 gpurun_out/r2_2gpu_group_tests.log; tail -3 gpurun_out/r2_2gpu_bench.err
