#!/bin/bash
# session 5, last call: the default bench line on the final build, then the ncu launch list of the bench command's main leg
mkdir -p gpurun_out
timeout 200 python bench.py > gpurun_out/r5c_bench.json 2> gpurun_out/r5c_bench.err; echo "bench rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-ode --no-train --no-extras --no-config5 --no-cpu-baseline"
timeout 100 $BENCH > gpurun_out/r5c_bench_plain.json 2> gpurun_out/r5c_bench_plain.err &&
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r5c_bench_launches.csv $BENCH > gpurun_out/r5c_ncu.log 2>&1
echo "bench launch list rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/r5c_bench.json').read().strip().splitlines()[-1])
print({k:v for k,v in d.items() if isinstance(v,(int,float))}, d['clocks'], d['dropin_predict_batch']['value'])"
