#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_config2.py > gpurun_out/r3p_sweep2.log 2>&1; echo "sweep2 rc=$?"; cp gpurun_out/sweep_config2.json gpurun_out/r3_sweep_config2.json
tail -20 gpurun_out/r3p_sweep2.log | cut -c1-220
