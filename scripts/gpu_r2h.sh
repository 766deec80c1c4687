#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lstm.py -m gpu -x -q -s -k "tensorcore_recurrence or chunked or fp32_forward" > gpurun_out/r2k_unit.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_unit.log
tail -22 gpurun_out/r2k_unit.log
timeout 300 python scripts/time_fp32_tc.py > gpurun_out/r2k_time_fp32.log 2>&1; tail -8 gpurun_out/r2k_time_fp32.log
