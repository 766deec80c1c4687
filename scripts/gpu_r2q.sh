#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_preproc.py -m gpu -x -q > gpurun_out/r2q_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2q_pytest.log; tail -6 gpurun_out/r2q_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2q_smoke.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-ode --no-train --no-extras --no-cpu-baseline > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print(d['config5']); print(d['e2e']['value'], d['value'])"
