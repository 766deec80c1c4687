#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3ag_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r3ag_tests.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r3ag_bench.json 2> gpurun_out/r3ag_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/r3ag_bench.json
