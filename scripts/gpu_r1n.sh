#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_n.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_n.log
timeout 600 python bench.py > gpurun_out/bench_n.json 2> gpurun_out/bench_n.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_n.json
python scripts/prof_fwd.py 16896 > gpurun_out/prof_plain_g.log 2>&1 || { echo "plain failed"; tail -3 gpurun_out/prof_plain_g.log; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1g.csv python scripts/prof_fwd.py 16896 > gpurun_out/ncu_launch_g.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"attn_pool_stream_bf16|head_mlp_kernel" -s 2 -c 2 -f -o gpurun_out/prof_pool_r1g python scripts/prof_fwd.py 16896 > gpurun_out/ncu_full_g.log 2>&1; echo "full rc=$?"
ls -la gpurun_out | tail -8
