#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_r2_g.sh
bash tools/gpu_r2_h.sh
