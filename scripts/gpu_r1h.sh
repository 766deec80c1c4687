#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py tests/test_gpu_ablation.py tests/test_gpu_next_rows.py tests/test_gpu_tensorcore.py tests/test_gpu_lstm.py tests/test_gpu_ode.py -q > gpurun_out/pytest_gpu_h.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu_h.log
timeout 200 python scripts/time_fp32.py 128 2>&1 | tail -2
