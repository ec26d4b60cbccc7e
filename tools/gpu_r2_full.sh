#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2full_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2full_status.txt
timeout 600 python bench.py --no-fit > gpurun_out/r2full_bench.json 2> gpurun_out/r2full_bench.err; echo "bench rc=$?" >> gpurun_out/r2full_status.txt
cat gpurun_out/r2full_status.txt; tail -8 gpurun_out/r2full_tests.log | cut -c1-300; tail -3 gpurun_out/r2full_bench.err
