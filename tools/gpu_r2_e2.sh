#!/bin/bash
bash tools/gpu_r2_e.sh
python bench.py --steps 5 --warmup 3 --no-fit --no-cpu --no-extra > gpurun_out/r2e_bench.json 2>gpurun_out/r2e_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/r2e_bench.json').read().strip().splitlines()[-1]); print(d['value'], d['roofline']['step_breakdown_ms'], d['roofline']['peaks'])"
