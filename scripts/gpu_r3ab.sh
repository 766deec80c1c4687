#!/bin/bash
mkdir -p gpurun_out
python scripts/train_step_once.py mixed 3 128 > gpurun_out/r3ab_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3ab_launches.csv python scripts/train_step_once.py mixed 3 128 > gpurun_out/r3ab_ncu.log 2>&1; echo "ncu rc=$?"
