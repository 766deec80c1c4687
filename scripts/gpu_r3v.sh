#!/bin/bash
mkdir -p gpurun_out
python scripts/time_gemm_single.py > gpurun_out/r3v_pair.log 2>&1; cat gpurun_out/r3v_pair.log
BCI_GEMM_PAIR=off python scripts/time_gemm_single.py > gpurun_out/r3v_1sm.log 2>&1; cat gpurun_out/r3v_1sm.log
