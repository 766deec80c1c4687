#!/bin/bash
# round 2, third session: ncu evidence for the CTA-pair bf16 projection GEMM (H = 256) and the launch list of the bench command
mkdir -p gpurun_out
python scripts/prof_fwd256.py > gpurun_out/r4d_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r4d_h256_launches.csv python scripts/prof_fwd256.py > gpurun_out/r4d_ncu1.log 2>&1
echo "h256 launch list rc=$?"
python scripts/prof_fwd256.py > gpurun_out/r4d_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"proj_gemm_bf16_pair" -s 4 -c 3 -o gpurun_out/r4d_pair_gemm python scripts/prof_fwd256.py > gpurun_out/r4d_ncu2.log 2>&1
echo "pair gemm full rc=$?"
BENCH="python bench.py --steps 2 --warmup 3 --no-ode --no-train --no-extras --no-config5 --no-cpu-baseline"
$BENCH > gpurun_out/r4d_bench_plain.json 2> gpurun_out/r4d_bench_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r4d_bench_launches.csv $BENCH > gpurun_out/r4d_ncu3.log 2>&1
echo "bench launch list rc=$?"
ls -la gpurun_out | grep r4d
