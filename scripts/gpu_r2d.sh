#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/prof_fp32_tc.py 2 > gpurun_out/r2d_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:lstm_rec_f16x3 -s 3 -c 1 -o gpurun_out/prof_fp32tc_r2d -f python scripts/prof_fp32_tc.py 2 > gpurun_out/r2d_ncu.log 2>&1
tail -3 gpurun_out/r2d_ncu.log
