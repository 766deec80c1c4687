#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scripts/tf32x3_check.py > gpurun_out/tf32x3_check.log 2>&1; echo "check rc=$?"; grep -v "explicit_hi=1" gpurun_out/tf32x3_check.log | tail -12
timeout 600 python -m pytest tests/test_gpu_lstm.py tests/test_gpu_train.py tests/test_gpu_ablation.py tests/test_gpu_next_rows.py tests/test_gpu_ode.py -x -q > gpurun_out/pytest_gpu_g.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu_g.log
for H in 128 256; do
  timeout 200 python scripts/time_fp32.py $H 2>&1 | tail -2
  BCI_FP32_GEMM=simt timeout 200 python scripts/time_fp32.py $H 2>&1 | tail -2
done
