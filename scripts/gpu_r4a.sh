#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pipeline.py -x -q > gpurun_out/r4a_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r4a_tests.log
{
for thr in 8 16; do for bf in 1 0; do
  echo "== BCI_STAGING_THREADS=$thr BCI_STAGING_BF16=$bf"
  BCI_STAGING_THREADS=$thr BCI_STAGING_BF16=$bf timeout 300 python scripts/time_dropin.py 67584
done; done
} > gpurun_out/r4a_dropin.log 2>&1
cat gpurun_out/r4a_dropin.log
{
echo "== pair"; timeout 300 python scripts/time_fp32_tc.py
echo "== BCI_GEMM_PAIR=off"; BCI_GEMM_PAIR=off timeout 300 python scripts/time_fp32_tc.py
} > gpurun_out/r4a_fp32.log 2>&1
cat gpurun_out/r4a_fp32.log | cut -c1-250
