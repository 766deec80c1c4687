#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3o_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3o_tests.log | cut -c1-250
timeout 300 python scripts/time_fp32_small.py bf16 > gpurun_out/r3o_small_bf16.log 2>&1; cat gpurun_out/r3o_small_bf16.log
BCI_BF16_SMALL=off timeout 300 python scripts/time_fp32_small.py bf16 > gpurun_out/r3o_small_bf16_old.log 2>&1; cat gpurun_out/r3o_small_bf16_old.log
