#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pipeline.py -q > gpurun_out/pytest_gpu_multi_k.log 2>&1; echo "pytest multi rc=$?"; tail -3 gpurun_out/pytest_gpu_multi_k.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_k.json 2> gpurun_out/bench_n2_k.err; echo "bench n2 rc=$?"; cut -c1-200 gpurun_out/bench_n2_k.json
