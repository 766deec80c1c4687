#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -m gpu -q --timeout 300 -s > gpurun_out/pytest_tc.log 2>&1
echo "pytest tc exit $?" >> gpurun_out/pytest_tc.log
grep -E "passed|failed|bf16 vs|Error|error|assert" gpurun_out/pytest_tc.log | head -40
timeout 900 python -m pytest tests -m gpu -q --timeout 600 --deselect tests/test_gpu_tensorcore.py > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
timeout 600 python scripts/quick_time.py bf16 > gpurun_out/quick_time_bf16.log 2>&1
echo "quick_time exit $?" >> gpurun_out/quick_time_bf16.log
grep -v "^ode" gpurun_out/quick_time_bf16.log
