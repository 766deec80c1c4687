#!/bin/bash
mkdir -p gpurun_out
python scripts/prof_train.py > gpurun_out/prof_train_plain_f.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_train_plain_f.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_train_f.csv python scripts/prof_train.py > gpurun_out/ncu_train_f.log 2>&1
echo "launch list rc=$?"
