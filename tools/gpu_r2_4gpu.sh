#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_4gpu_bench.json 2> gpurun_out/r2_4gpu_bench.err; echo "bench rc=$?" > gpurun_out/r2_4gpu_status.txt
cat gpurun_out/r2_4gpu_status.txt; tail -2 gpurun_out/r2_4gpu_bench.err
