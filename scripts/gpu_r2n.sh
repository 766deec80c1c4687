#!/bin/bash
# round 2 profiles: launch list of the bench command + full captures of the dominant kernels (each only after its plain run passed)
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-ode --no-train --no-extras --no-config5 --no-cpu-baseline"
timeout 300 $B > gpurun_out/r2n_bench_plain.json 2> gpurun_out/r2n_bench_plain.err || { tail -5 gpurun_out/r2n_bench_plain.err; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2n_launches.csv $B > gpurun_out/r2n_ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 120 python scripts/prof_fwd.py 16896 > gpurun_out/r2n_fwd_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_fused_bf16" -s 3 -c 3 -f -o gpurun_out/prof_k23_r2n python scripts/prof_fwd.py 16896 > gpurun_out/r2n_ncu_k23.log 2>&1; echo "k23 rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"input_proj_bf16|attn_pool_stream_bf16|head_mlp" -s 3 -c 3 -f -o gpurun_out/prof_k1pool_r2n python scripts/prof_fwd.py 16896 > gpurun_out/r2n_ncu_k1pool.log 2>&1; echo "k1/pool rc=$?"
timeout 120 python scripts/prof_ode.py > gpurun_out/r2n_ode_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"ode_rk4x2" -s 1 -c 1 -f -o gpurun_out/prof_ode_r2n python scripts/prof_ode.py > gpurun_out/r2n_ncu_ode.log 2>&1; echo "ode rc=$?"
timeout 120 python scripts/prof_fp32_tc.py 2 > gpurun_out/r2n_fp32_plain.log 2>&1 || exit 1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"lstm_rec_f16x3|gemm_tf32x3_kernel" -s 7 -c 7 -f -o gpurun_out/prof_fp32tc_r2n python scripts/prof_fp32_tc.py 2 > gpurun_out/r2n_ncu_fp32.log 2>&1; echo "fp32 rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2n_launches_fp32.csv python scripts/prof_fp32_tc.py 1 > /dev/null 2>&1; echo "fp32 launch list rc=$?"
ls -la gpurun_out | grep r2n
