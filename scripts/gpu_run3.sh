#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 300 -s 2>&1 | grep -E "passed|failed|bf16 vs|Error|assert" | head -20
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-ode > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo rc=$?; tail -3 gpurun_out/bench2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench2.json'))
print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['phase_ms_per_step'], d['roofline']['whole_path'])
PY
