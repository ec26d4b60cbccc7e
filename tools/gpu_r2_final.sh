#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2final_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2final_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2final_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2final_status.txt
timeout 1200 python bench.py > gpurun_out/r2final_bench.json 2> gpurun_out/r2final_bench.err; echo "bench rc=$?" >> gpurun_out/r2final_status.txt
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2final_ref.json 2> gpurun_out/r2final_ref.err; echo "ref rc=$?" >> gpurun_out/r2final_status.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra --no-fit > gpurun_out/r2final_ncu.log 2>&1; echo "ncu rc=$?" >> gpurun_out/r2final_status.txt
cat gpurun_out/r2final_status.txt; tail -n 6 gpurun_out/r2final_tests.log | cut -c1-300; tail -n 2 gpurun_out/r2final_smoke.log; tail -n 3 gpurun_out/r2final_bench.err
