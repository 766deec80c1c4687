#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ablation.py tests/test_gpu_next_rows.py -q > gpurun_out/pytest_gpu_i.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu_i.log
timeout 200 python scripts/time_fp32.py 128 2>&1 | tail -2
