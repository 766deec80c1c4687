#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3m_bench_2gpu.json 2> gpurun_out/r3m_bench_2gpu.err; echo "bench rc=$?"
tail -c 2500 gpurun_out/r3m_bench_2gpu.json; tail -5 gpurun_out/r3m_bench_2gpu.err
