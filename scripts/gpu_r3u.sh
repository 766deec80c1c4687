#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_tensorcore.py -x -q -s -k "single_pass" > gpurun_out/r3w_tests.log 2>&1; echo "tests rc=$?"
grep -E "single-pass|passed|failed|Error|error" gpurun_out/r3w_tests.log | cut -c1-220 | tail -12
timeout 600 python -m pytest tests/test_gpu_rec_swap.py tests/test_gpu_train.py -x -q -k "not jitter" > gpurun_out/r3w_tests2.log 2>&1; echo "tests2 rc=$?"
tail -3 gpurun_out/r3w_tests2.log | cut -c1-200
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3w_time.log 2>&1; echo "time rc=$?"
head -4 gpurun_out/r3w_time.log | cut -c1-200
