#!/bin/bash
mkdir -p gpurun_out
python scripts/train_step_once.py mixed 3 256 > gpurun_out/r3l_plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3_train_mixed_h256_launches.csv python scripts/train_step_once.py mixed 3 256 > gpurun_out/r3l_ncu1.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lstm_rec_swap256_fwd|lstm_bptt_swap256' -s 8 -c 2 -o gpurun_out/r3_swap256 python scripts/train_step_once.py mixed 3 256 > gpurun_out/r3l_ncu2.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lstm_bptt_swap' -s 6 -c 1 -o gpurun_out/r3_bptt_mixed python scripts/train_step_once.py mixed 3 > gpurun_out/r3l_ncu3.log 2>&1; echo "ncu full rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lstm_bptt_swap' -s 6 -c 1 -o gpurun_out/r3_bptt_fp32 python scripts/train_step_once.py fp32 3 > gpurun_out/r3l_ncu4.log 2>&1; echo "ncu full rc=$?"
tail -2 gpurun_out/r3l_plain.log
