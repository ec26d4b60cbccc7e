#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extra --fit-iterations 1 > gpurun_out/r2d_2gpu_bench.json 2> gpurun_out/r2d_2gpu_bench.err; echo "bench rc=$?" > gpurun_out/r2d_2gpu_status.txt
timeout 300 python -m pytest tests/test_gpu_group.py -m gpu -x -q -k "fit_through" > gpurun_out/r2d_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/r2d_2gpu_status.txt
cat gpurun_out/r2d_2gpu_status.txt; tail -n 3 gpurun_out/r2d_2gpu_bench.err; tail -n 5 gpurun_out/r2d_tests.log
