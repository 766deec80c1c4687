#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_fwd.py 9472 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python scripts/prof_fwd.py 9472 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python scripts/prof_fwd.py 9472 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lstm_rec_bf16|proj_gemm_bf16" -s 8 -c 2 -o gpurun_out/prof_rec_gemm_r1 python scripts/prof_fwd.py 9472 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
