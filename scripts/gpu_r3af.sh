#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_train.py -x -q -k "dropout" > gpurun_out/r3af_tests.log 2>&1; echo "tests rc=$?"
tail -6 gpurun_out/r3af_tests.log | cut -c1-300
