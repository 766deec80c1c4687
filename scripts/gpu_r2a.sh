#!/bin/bash
# round 2, first GPU pass: full GPU test suite, printed errors of the tightened bf16 tests, one bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv,noheader > gpurun_out/r2s_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2s_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2s_pytest.log
timeout 600 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_train.py tests/test_gpu_next_rows.py -m gpu -q -s -k "oracle or full_wave or config3 or seeded" > gpurun_out/r2s_pytest_s.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err; echo "bench rc=$?" >> gpurun_out/r2s_bench.err
tail -5 gpurun_out/r2s_pytest.log; tail -3 gpurun_out/r2s_bench.err; head -c 1500 gpurun_out/r2s_bench.json
