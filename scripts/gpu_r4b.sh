#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4b_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/r4b_tests.log; echo "tests took $SECONDS s"; SECONDS=0
timeout 900 python bench.py > gpurun_out/r4b_bench.json 2> gpurun_out/r4b_bench.err; echo "bench rc=$?"
echo "bench took $SECONDS s"
tail -c 1500 gpurun_out/r4b_bench.json
