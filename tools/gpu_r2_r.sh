#!/bin/bash
mkdir -p gpurun_out
for lc in 64 96 128 192 256 384 512; do echo "LC=$lc"; GPAR_SH_LC=$lc timeout 300 python tools/prof_smooth_shared_one.py 2>&1 | grep smooth | tail -1; done > gpurun_out/r2r_lc.txt
cat gpurun_out/r2r_lc.txt
