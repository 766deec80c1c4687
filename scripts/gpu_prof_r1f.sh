#!/bin/bash
# capture F: training step (512 windows, fp32) with the split-precision tcgen05 GEMMs and the resident-W_hh recurrences
mkdir -p gpurun_out
python scripts/prof_train.py > gpurun_out/prof_plain_f.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/prof_plain_f.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1f.csv python scripts/prof_train.py > gpurun_out/ncu_launch_f.log 2>&1
echo "launch list rc=$?"
# second step: the 22 GEMM launches of step 1 are skipped, then 11 captured (3 forward NT, then the top layer's TN / NT and the next one's)
timeout 600 ncu --set full --clock-control none -k regex:"gemm_tf32x3_kernel" -s 22 -c 11 -f -o gpurun_out/prof_tf32x3_r1f python scripts/prof_train.py > gpurun_out/ncu_full_f1.log 2>&1
echo "full gemm rc=$?"
timeout 600 ncu --set full --clock-control none -k regex:"lstm_rec_f32|lstm_bptt_f32" -s 6 -c 4 -f -o gpurun_out/prof_rec_r1f python scripts/prof_train.py > gpurun_out/ncu_full_f2.log 2>&1
echo "full rec rc=$?"
ls -la gpurun_out
