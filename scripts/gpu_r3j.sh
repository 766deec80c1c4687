#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r3j_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r3j_tests.log | cut -c1-250
timeout 300 python scripts/time_fp32_small.py > gpurun_out/r3j_small_tc.log 2>&1; cat gpurun_out/r3j_small_tc.log
BCI_TRAIN_REC=simt timeout 300 python scripts/time_fp32_small.py > gpurun_out/r3j_small_simt.log 2>&1; cat gpurun_out/r3j_small_simt.log
