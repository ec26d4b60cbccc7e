#!/bin/bash
mkdir -p gpurun_out
python tools/standin_reference_outputs.py gpurun_out/standin_ref.json > gpurun_out/r2b_status.txt 2>&1
GPAR_REFERENCE_GOLDEN=gpurun_out/standin_ref.json timeout 600 python -m pytest tests/test_reference_golden.py -m gpu -x -q > gpurun_out/r2b_golden_gpu.log 2>&1; echo "golden-gpu rc=$?" >> gpurun_out/r2b_status.txt
timeout 900 python tools/kalman_sweep.py > gpurun_out/r2b_kalman_sweep.txt 2>&1; echo "sweep rc=$?" >> gpurun_out/r2b_status.txt
cat gpurun_out/r2b_status.txt; tail -5 gpurun_out/r2b_golden_gpu.log; cat gpurun_out/r2b_kalman_sweep.txt
