#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_tensorcore.py tests/test_gpu_lstm.py -x -q -s -k "staggered or full_chunk or full_pass" > gpurun_out/r4i_tests.log 2>&1; rc=$?; echo "tests rc=$rc"
grep -E "passed|failed|oracle|Error" gpurun_out/r4i_tests.log | tail -12 | cut -c1-250
if [ $rc -ne 0 ]; then tail -30 gpurun_out/r4i_tests.log | cut -c1-300; exit 0; fi
timeout 120 python scripts/prof_fwd256.py > gpurun_out/r4i_plain.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lstm_rec256_bf16_pipe" -s 2 -c 2 -o gpurun_out/r4i_rec256_pipe python scripts/prof_fwd256.py > gpurun_out/r4i_ncu1.log 2>&1
echo "rec256 pipe ncu rc=$?"
timeout 120 python scripts/prof_fp32_tc.py > gpurun_out/r4i_plain2.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lstm_rec_f16x3_pipe" -s 3 -c 2 -o gpurun_out/r4i_f16x3_pipe python scripts/prof_fp32_tc.py > gpurun_out/r4i_ncu2.log 2>&1
echo "f16x3 pipe ncu rc=$?"
ls -la gpurun_out | grep r4i
