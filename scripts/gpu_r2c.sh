#!/bin/bash
# round 2: packed ODE kernel + fp32 tensor-core recurrence: unit tests first, then timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_ode.py tests/test_gpu_lstm.py -m gpu -x -q -s -k "f16x3 or packed or tensorcore_recurrence or chunked" > gpurun_out/r2c_unit.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_unit.log
tail -25 gpurun_out/r2c_unit.log
timeout 300 python scripts/time_fp32_tc.py > gpurun_out/r2c_time_fp32.log 2>&1; BCI_FP32_REC=simt timeout 300 python scripts/time_fp32_tc.py >> gpurun_out/r2c_time_fp32.log 2>&1; tail -16 gpurun_out/r2c_time_fp32.log
timeout 300 python scripts/ode_time.py > gpurun_out/r2c_ode_time.log 2>&1; BCI_ODE_RK4=scalar timeout 300 python scripts/ode_time.py >> gpurun_out/r2c_ode_time.log 2>&1; tail -16 gpurun_out/r2c_ode_time.log
