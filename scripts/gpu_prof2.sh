#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
python scripts/prof_fwd.py 9472 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python scripts/prof_fwd.py 9472 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python scripts/prof_fwd.py 9472 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lstm_rec_bf16|proj_gemm_bf16|input_proj_bf16|attn_score_bf16|attn_pool_finish" -s 14 -c 9 -o gpurun_out/prof_bf16_r1 python scripts/prof_fwd.py 9472 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full.log
