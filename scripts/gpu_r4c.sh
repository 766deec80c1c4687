#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -k "proj_gemm or h256" > gpurun_out/r4c_tests.log 2>&1; echo "tests rc=$?"
tail -15 gpurun_out/r4c_tests.log | cut -c1-300
{
echo "== pair"; timeout 300 python scripts/time_h256.py
echo "== BCI_GEMM_PAIR=off"; BCI_GEMM_PAIR=off timeout 300 python scripts/time_h256.py
} > gpurun_out/r4c_h256.log 2>&1
cat gpurun_out/r4c_h256.log | cut -c1-250
