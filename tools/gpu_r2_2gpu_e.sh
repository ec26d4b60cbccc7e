#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/check_row_sharded_torchrun.py 200003 300 2 > gpurun_out/r2e_small.json 2> gpurun_out/r2e_small.err; echo "small rc=$?" > gpurun_out/r2e_status.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 tools/check_row_sharded_torchrun.py > gpurun_out/r2e_cfg5.json 2> gpurun_out/r2e_cfg5.err; echo "cfg5 rc=$?" >> gpurun_out/r2e_status.txt
cat gpurun_out/r2e_status.txt; cat gpurun_out/r2e_small.json gpurun_out/r2e_cfg5.json; tail -n 5 gpurun_out/r2e_small.err; tail -n 5 gpurun_out/r2e_cfg5.err
