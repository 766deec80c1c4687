#!/bin/bash
mkdir -p gpurun_out
for mode in f32 f16; do
  echo "== BCI_REC_ACT=$mode"
  BCI_REC_ACT=$mode timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 300 -s 2>&1 | grep -E "passed|failed|bf16 vs|Error|assert" | head -20
  BCI_REC_ACT=$mode timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ode --no-train > gpurun_out/bench_$mode.json 2> gpurun_out/bench_$mode.err; echo rc=$?; tail -3 gpurun_out/bench_$mode.err
  python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/bench_$mode.json') if l.startswith('{')][-1])
print(d['value'], d['ms_per_step'], d['roofline']['phase_ms_per_step'])
PY
done
