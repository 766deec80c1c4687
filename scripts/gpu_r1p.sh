#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ablation.py tests/test_gpu_next_rows.py -q 2>&1 | tail -2
timeout 200 python scripts/time_fp32.py 128 2>&1 | tail -2
timeout 600 python bench.py --no-ode --no-extras > gpurun_out/bench_p.json 2> gpurun_out/bench_p.err; echo "bench rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_p.json').read().strip().splitlines()[-1]); print(d['value'], d['clocks'], d['train_step']['ms_per_step'])"
