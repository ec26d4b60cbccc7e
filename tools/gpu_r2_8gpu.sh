#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | wc -l > gpurun_out/r2_8gpu_status.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu > gpurun_out/r2_8gpu_bench.json 2> gpurun_out/r2_8gpu_bench.err; echo "bench rc=$?" >> gpurun_out/r2_8gpu_status.txt
cat gpurun_out/r2_8gpu_status.txt; tail -3 gpurun_out/r2_8gpu_bench.err
