#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tensorcore.py -x -q -s -k "gemm" > gpurun_out/r3ak_tests.log 2>&1; echo "tests rc=$?"
grep -E "GEMM|gemm|passed|failed|Error|error" gpurun_out/r3ak_tests.log | cut -c1-200 | tail -30
timeout 900 python -m pytest tests/test_gpu_rec_swap.py tests/test_gpu_train.py tests/test_gpu_lstm.py tests/test_gpu_ablation.py tests/test_gpu_next_rows.py -x -q -s -k "not jitter" > gpurun_out/r3ak_tests2.log 2>&1; echo "tests2 rc=$?"
grep -E "mixed step|passed|failed" gpurun_out/r3ak_tests2.log | cut -c1-200 | tail -8
timeout 300 python scripts/time_train_modes.py 10 > gpurun_out/r3ak_time.log 2>&1; echo "time rc=$?"
head -4 gpurun_out/r3ak_time.log | cut -c1-200
