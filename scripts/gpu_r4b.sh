#!/bin/bash
mkdir -p gpurun_out
{
for bf in 1 0; do
  echo "== threads auto BCI_STAGING_BF16=$bf"
  BCI_STAGING_BF16=$bf timeout 300 python scripts/time_dropin.py 67584
done
echo "== threads 12"; BCI_STAGING_THREADS=12 timeout 300 python scripts/time_dropin.py 67584
echo "== fp32 engine"; timeout 300 python scripts/time_dropin.py 33792 fp32
} > gpurun_out/r4b_dropin.log 2>&1
cat gpurun_out/r4b_dropin.log
/usr/bin/time -v timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r4b_tests.log 2>&1; echo "tests rc=$?"
grep -E "passed|failed|Elapsed" gpurun_out/r4b_tests.log
/usr/bin/time -v timeout 900 python bench.py > gpurun_out/r4b_bench.json 2> gpurun_out/r4b_bench.err; echo "bench rc=$?"
grep -E "Elapsed" gpurun_out/r4b_bench.err
tail -c 1500 gpurun_out/r4b_bench.json
