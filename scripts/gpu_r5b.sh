#!/bin/bash
# session 5: fused-kernel epilogue variants -- bf16 parity / jitter tests, then two short main-leg bench runs
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -x -q -m gpu 2>&1 | tail -2
for i in 1 2; do
  timeout 120 python bench.py --steps 30 --warmup 3 --no-extras --no-train --no-config5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['value']), round(d['ms_per_step'],3), d['clocks']['sm_mhz'], d['roofline']['phase_ms_per_step'], round(d['e2e']['value']))"
done
