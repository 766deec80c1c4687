#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1 2; do echo "== BCI_TC_PF=$pf"; BCI_TC_PF=$pf timeout 300 python scripts/time_fp32_tc.py 2>&1 | grep -E "B=  4096|B=  9472"; done > gpurun_out/r2g_pf.log 2>&1
cat gpurun_out/r2g_pf.log
