#!/bin/bash
# round 2, call A: GPU parity tests, short bench (fit with a tiny budget), scaled-gradient phase timings + ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2a_status.txt
timeout 300 python tools/prof_scaled_grad.py > gpurun_out/r2a_scaled_grad.txt 2>&1; echo "prof rc=$?" >> gpurun_out/r2a_status.txt
timeout 900 python bench.py --steps 5 --warmup 3 --fit-iterations 1 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches_scaled_grad.csv python tools/prof_scaled_grad.py > gpurun_out/r2a_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2a_status.txt
cat gpurun_out/r2a_status.txt; tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_scaled_grad.txt
