#!/bin/bash
# launch lists of the config-3 step in both modes + one full ncu capture of the four swapped recurrence kernels
mkdir -p gpurun_out
python scripts/train_step_once.py mixed 3 > gpurun_out/r3h_plain_mixed.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3_train_mixed_launches.csv python scripts/train_step_once.py mixed 3 > gpurun_out/r3h_ncu1.log 2>&1; echo "ncu mixed rc=$?"
python scripts/train_step_once.py fp32 3 > gpurun_out/r3h_plain_fp32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3_train_fp32_launches.csv python scripts/train_step_once.py fp32 3 > gpurun_out/r3h_ncu2.log 2>&1; echo "ncu fp32 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lstm_rec_swap_fwd|lstm_bptt_swap' -s 6 -c 2 -o gpurun_out/r3_swap_mixed python scripts/train_step_once.py mixed 3 > gpurun_out/r3h_ncu3.log 2>&1; echo "ncu full mixed rc=$?"
ncu --set full --clock-control none --import-source on -k regex:'lstm_rec_swap_fwd|lstm_bptt_swap' -s 6 -c 2 -o gpurun_out/r3_swap_fp32 python scripts/train_step_once.py fp32 3 > gpurun_out/r3h_ncu4.log 2>&1; echo "ncu full fp32 rc=$?"
ls -la gpurun_out/*.ncu-rep
