#!/bin/bash
mkdir -p gpurun_out
SECONDS=0; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4g_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed" gpurun_out/r4g_tests.log; echo "tests took $SECONDS s"; SECONDS=0
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r4g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r4g_smoke.log
SECONDS=0
timeout 600 python bench.py > gpurun_out/r4g_bench.json 2> gpurun_out/r4g_bench.err; echo "bench rc=$?"
echo "bench took $SECONDS s"
tail -c 1200 gpurun_out/r4g_bench.json
