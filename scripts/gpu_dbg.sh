#!/bin/bash
for d in 0 1 2 3 4 7; do
BCI_REC_DBG=$d python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-ode --no-train > gpurun_out/b.json 2>/dev/null
python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/b.json') if l.startswith('{')][-1])
print("dbg=$d rec ms", round(d['roofline']['phase_ms_per_step']['recurrence'],3))
PY
done
