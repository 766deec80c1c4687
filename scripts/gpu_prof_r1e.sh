#!/bin/bash
# launch list + ncu --set full of the H = 256 bf16 kernels (after a plain run exited 0)
mkdir -p gpurun_out
python scripts/prof_fwd256.py 8448 > gpurun_out/prof_plain_e.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_e.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1e.csv python scripts/prof_fwd256.py 8448 > gpurun_out/ncu_launch_e.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"lstm_rec256_bf16|proj_gemm_bf16|attn_score256|ln_gelu_rows256|x_to_bf16" -s 9 -c 7 -f -o gpurun_out/prof_bf16_r1e python scripts/prof_fwd256.py 8448 > gpurun_out/ncu_full_e.log 2>&1
echo "full rc=$?"
tail -2 gpurun_out/ncu_full_e.log
