#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_fwd.py 9472 > gpurun_out/prof_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1c.csv python scripts/prof_fwd.py 9472 > gpurun_out/ncu_launch.log 2>&1
echo "launch list rc=$?"
python scripts/prof_fwd.py 9472 > gpurun_out/prof_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"lstm_rec_bf16|proj_gemm_bf16|input_proj_bf16|attn_score_bf16|attn_pool_finish" -s 17 -c 8 -o gpurun_out/prof_bf16_r1c python scripts/prof_fwd.py 9472 > gpurun_out/ncu_full.log 2>&1
echo "full rc=$?"
python scripts/ode_time.py > gpurun_out/ode_time.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"ode_rk4_kernel" -s 6 -c 1 -o gpurun_out/prof_ode_r1 python scripts/ode_time.py > gpurun_out/ncu_ode.log 2>&1
echo "ode rc=$?"
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_final.json
