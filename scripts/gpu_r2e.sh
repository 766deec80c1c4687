#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lstm.py -m gpu -x -q -s -k "f16x3 or tensorcore_recurrence or chunked" > gpurun_out/r2f_unit.log 2>&1; echo "rc=$?" >> gpurun_out/r2f_unit.log
tail -14 gpurun_out/r2f_unit.log
timeout 300 python scripts/time_fp32_tc.py > gpurun_out/r2f_time_fp32.log 2>&1; tail -8 gpurun_out/r2f_time_fp32.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_f16x3 -s 3 -c 1 -o gpurun_out/prof_fp32tc_r2f -f python scripts/prof_fp32_tc.py 2 > gpurun_out/r2f_ncu.log 2>&1
tail -2 gpurun_out/r2f_ncu.log
