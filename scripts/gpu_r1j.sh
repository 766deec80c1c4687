#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_j.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_j.log
timeout 200 python scripts/time_fp32.py 128 2>&1 | tail -2
timeout 200 python scripts/time_fp32.py 256 2>&1 | tail -2
timeout 300 python scripts/torch_cuda_baseline.py 128 > gpurun_out/torch_cuda_h128.json 2> gpurun_out/torch_cuda_h128.err; echo "torch128 rc=$?"
timeout 300 python scripts/torch_cuda_baseline.py 256 > gpurun_out/torch_cuda_h256.json 2> gpurun_out/torch_cuda_h256.err; echo "torch256 rc=$?"
timeout 600 python bench.py > gpurun_out/bench_j.json 2> gpurun_out/bench_j.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_j.json
