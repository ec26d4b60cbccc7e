#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2j_status.txt
timeout 1200 python bench.py > gpurun_out/r2j_bench.json 2>gpurun_out/r2j_bench.err; echo "bench rc=$?" >> gpurun_out/r2j_status.txt
cat gpurun_out/r2j_status.txt; tail -5 gpurun_out/r2j_tests.log; tail -3 gpurun_out/r2j_bench.err
