#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/kalman_grad_time.py > gpurun_out/r2m_grad.txt 2>&1; echo "grad rc=$?" > gpurun_out/r2m_status.txt
cat gpurun_out/r2m_status.txt; cat gpurun_out/r2m_grad.txt
