#!/bin/bash
mkdir -p gpurun_out
{ lscpu | grep -E "^CPU\(s\)|NUMA|Socket"; nvidia-smi topo -m | head -12; grep MemTotal /proc/meminfo; } > gpurun_out/r2p_topo8.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2p_bench_n8.json 2> gpurun_out/r2p_bench_n8.err; echo "bench rc=$?" >> gpurun_out/r2p_bench_n8.err
tail -3 gpurun_out/r2p_bench_n8.err; tail -c 1800 gpurun_out/r2p_bench_n8.json
