#!/bin/bash
# final launch lists of the training steps (H = 128 both modes, H = 256 mixed) + full ncu of the pair GEMM and the H = 256 pair recurrences
mkdir -p gpurun_out
for cfg in "mixed 3 128" "fp32 3 128" "mixed 3 256"; do
  set -- $cfg
  python scripts/train_step_once.py $1 $2 $3 > gpurun_out/r3y_plain_$1_$3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r3_train_$1_h$3_launches.csv python scripts/train_step_once.py $1 $2 $3 > gpurun_out/r3y_ncu_$1_$3.log 2>&1; echo "ncu $cfg rc=$?"
done
ncu --set full --clock-control none --import-source on -k regex:'gemm_tf32_pair_kernel' -s 8 -c 4 -o gpurun_out/r3_gemm_pair python scripts/train_step_once.py mixed 3 256 > gpurun_out/r3y_ncu_full.log 2>&1; echo "ncu full rc=$?"
