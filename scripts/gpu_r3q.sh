#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_rec_swap.py -x -q -k "ablation or dropout" > gpurun_out/r3q_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3q_tests.log | cut -c1-300
