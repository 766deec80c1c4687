#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py tests/test_gpu_train.py -x -q -s > gpurun_out/r3d_tests.log 2>&1; echo "tests rc=$?"
tail -12 gpurun_out/r3d_tests.log | cut -c1-250
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3d_time.log 2>&1; echo "time rc=$?"
tail -12 gpurun_out/r3d_time.log | cut -c1-200
python scripts/train_step_once.py mixed 3 > gpurun_out/r3d_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3d_launches_mixed.csv python scripts/train_step_once.py mixed 3 > gpurun_out/r3d_ncu.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r3d_plain.log
