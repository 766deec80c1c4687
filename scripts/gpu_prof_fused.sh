#!/bin/bash
# one ncu --set full capture of the fused cluster kernel (layer 1: 256-wide input), after a plain run exited 0
mkdir -p gpurun_out
python scripts/prof_fwd.py 8448 > gpurun_out/prof_plain_fused.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_fused.log; exit 1; }
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lstm_fused_bf16" -s 4 -c 1 -f -o gpurun_out/prof_fused_$1 python scripts/prof_fwd.py 8448 > gpurun_out/ncu_fused.log 2>&1
echo "full rc=$?"
tail -3 gpurun_out/ncu_fused.log
