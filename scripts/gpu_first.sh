#!/bin/bash
# first GPU contact: tests + a quick timing of each stage
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_time.py > gpurun_out/quick_time.log 2>&1
echo "quick_time exit $?" >> gpurun_out/quick_time.log
cat gpurun_out/quick_time.log
