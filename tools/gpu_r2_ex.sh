#!/bin/bash
mkdir -p gpurun_out
( python examples/gpar_scaled_example.py --iterations 150 | tail -1; python examples/gpar_scaled_example.py --iterations 150 | tail -1; python examples/gpar_scaled_example.py --iterations 150 --speculative | tail -1; python examples/gpar_scaled_example.py --iterations 150 --restarts 8 | tail -1 ) > gpurun_out/r2_example.txt 2>&1
cat gpurun_out/r2_example.txt
