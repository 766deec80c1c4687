#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py -x -q -s > gpurun_out/r3a_tests.log 2>&1; echo "tests rc=$?"
tail -40 gpurun_out/r3a_tests.log | cut -c1-250
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3a_time.log 2>&1; echo "time rc=$?"
tail -8 gpurun_out/r3a_time.log | cut -c1-200
