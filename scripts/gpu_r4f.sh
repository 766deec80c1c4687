#!/bin/bash
mkdir -p gpurun_out
timeout 90 python -m pytest tests/test_gpu_tensorcore.py -x -q -s -k "recurrence_h256" > gpurun_out/r4f_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
tail -8 gpurun_out/r4f_tests.log | cut -c1-250
if [ $rc -ne 0 ]; then exit 0; fi
timeout 240 python -m pytest tests/test_gpu_tensorcore.py -x -q -s -k "h256 or jitter" > gpurun_out/r4f_tests2.log 2>&1; rc=$?; echo "tests2 rc=$rc"
grep -E "passed|failed|oracle|Error" gpurun_out/r4f_tests2.log | tail -8 | cut -c1-250
if [ $rc -ne 0 ]; then exit 0; fi
{
echo "== pipe"; timeout 120 python scripts/time_h256.py
echo "== BCI_H256_PIPE=0"; BCI_H256_PIPE=0 timeout 120 python scripts/time_h256.py
} > gpurun_out/r4f_h256.log 2>&1
cat gpurun_out/r4f_h256.log | cut -c1-250
