#!/bin/bash
# session 5: gather kernel timing + ncu, smoke(), 2-GPU bench line is a separate call (gpu_final_n2.sh)
mkdir -p gpurun_out
timeout 120 python scripts/prof_permute.py > gpurun_out/r5_permute_timing.log 2>&1; echo "permute rc=$?"; cat gpurun_out/r5_permute_timing.log | tail -3
timeout 200 ncu --set full --clock-control none --import-source on -k regex:"permute_channels" -s 3 -c 1 -o gpurun_out/r5_permute python scripts/prof_permute.py > gpurun_out/r5_ncu_permute.log 2>&1; echo "ncu rc=$?"
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r5_smoke.log
