#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2last_tests.log 2>&1; echo "tests rc=$?" > gpurun_out/r2last_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2last_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2last_status.txt
cat gpurun_out/r2last_status.txt; tail -n 5 gpurun_out/r2last_tests.log | cut -c1-300; tail -n 1 gpurun_out/r2last_smoke.log
