#!/bin/bash
# launch list + one ncu --set full capture of every bf16 kernel of the fused path (after a plain run exited 0)
mkdir -p gpurun_out
python scripts/prof_fwd.py 16896 > gpurun_out/prof_plain_d.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_d.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1d.csv python scripts/prof_fwd.py 16896 > gpurun_out/ncu_launch_d.log 2>&1
echo "launch list rc=$?"
python scripts/prof_fwd.py 16896 > gpurun_out/prof_plain_d2.log 2>&1 &&
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_fused_bf16|input_proj_bf16|attn_score_bf16|attn_pool_finish" -s 5 -c 5 -f -o gpurun_out/prof_bf16_r1d python scripts/prof_fwd.py 16896 > gpurun_out/ncu_full_d.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full_d.log
