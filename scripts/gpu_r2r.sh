#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_preproc.py -m gpu -x -q > gpurun_out/r2r_pytest.log 2>&1; echo "rc=$?" >> gpurun_out/r2r_pytest.log; tail -4 gpurun_out/r2r_pytest.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-ode --no-train --no-cpu-baseline > gpurun_out/r2r_bench.json 2> gpurun_out/r2r_bench.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2r_bench.json')); print(d['config5']['value'], d['preprocess'])"
