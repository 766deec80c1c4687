#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py tests/test_gpu_train.py -x -q -s > gpurun_out/r3n_tests.log 2>&1; echo "tests rc=$?"
grep -E "swap|mixed step|passed|failed|Error|error|assert" gpurun_out/r3n_tests.log | cut -c1-220 | tail -30
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3n_time.log 2>&1; echo "time rc=$?"
tail -12 gpurun_out/r3n_time.log | cut -c1-200
