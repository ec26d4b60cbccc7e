#!/bin/bash
mkdir -p gpurun_out
GPAR_GROUP_LOOPBACK=1 timeout 300 python examples/sharded_output_example.py --devices 3 --same-device > gpurun_out/r2_sharded_example.txt 2>&1; echo "rc=$?" >> gpurun_out/r2_sharded_example.txt
timeout 300 python examples/gpar_scaled_example.py --iterations 60 2>&1 | tail -n 2 >> gpurun_out/r2_sharded_example.txt
tail -n 8 gpurun_out/r2_sharded_example.txt | cut -c1-300
