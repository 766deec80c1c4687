#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/sweep_config2.py > gpurun_out/r2t_sweep2.log 2>&1; echo "sweep2 rc=$?"; cp gpurun_out/sweep_config2.json gpurun_out/r2_sweep_config2.json
timeout 600 python scripts/sweep_config4.py > gpurun_out/r2t_sweep4.log 2>&1; echo "sweep4 rc=$?"; cp gpurun_out/sweep_config4.json gpurun_out/r2_sweep_config4.json
tail -15 gpurun_out/r2t_sweep2.log | cut -c1-200; tail -12 gpurun_out/r2t_sweep4.log | cut -c1-200
