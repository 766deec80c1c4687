#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py -x -q -k "jitter" > gpurun_out/r3r_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3r_tests.log | cut -c1-300
