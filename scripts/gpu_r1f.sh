#!/bin/bash
# re-verification of the whole GPU suite + library bar + bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_f.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu_f.log
timeout 300 python scripts/torch_cuda_baseline.py 128 > gpurun_out/torch_cuda_h128.json 2> gpurun_out/torch_cuda_h128.err; echo "torch128 rc=$?"; cat gpurun_out/torch_cuda_h128.json
timeout 300 python scripts/torch_cuda_baseline.py 256 > gpurun_out/torch_cuda_h256.json 2> gpurun_out/torch_cuda_h256.err; echo "torch256 rc=$?"; cat gpurun_out/torch_cuda_h256.json
timeout 600 python bench.py > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench_f.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
