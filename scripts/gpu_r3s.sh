#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py tests/test_gpu_train.py tests/test_gpu_tensorcore.py -x -q -k "not jitter" > gpurun_out/r3t_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3t_tests.log | cut -c1-300
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3t_time.log 2>&1; echo "time rc=$?"
head -4 gpurun_out/r3t_time.log | cut -c1-200
