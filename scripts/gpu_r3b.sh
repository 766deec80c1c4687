#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py -x -q -s > gpurun_out/r3c_tests.log 2>&1; echo "tests rc=$?"
tail -16 gpurun_out/r3c_tests.log | cut -c1-250
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3c_time_tmem.log 2>&1; echo "time rc=$?"
tail -12 gpurun_out/r3c_time_tmem.log | cut -c1-200
