#!/bin/bash
# round 2: compute-sanitizer runs on the cluster / mbarrier kernels (small sizes), logs kept under profiles/
mkdir -p gpurun_out
timeout 120 python scripts/sanitize.py all > gpurun_out/r2m_plain.log 2>&1 || { tail -5 gpurun_out/r2m_plain.log; exit 1; }
for tool in synccheck racecheck memcheck; do
  for w in fused pool h256 fp32tc; do
    echo "== $tool $w" >> gpurun_out/r2m_sanitizer.log
    timeout 240 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize.py $w 2>&1 | grep -E "ok|ERROR SUMMARY|RACECHECK SUMMARY|Error|hazard|Race|Barrier|=========     at|Invalid" | head -12 >> gpurun_out/r2m_sanitizer.log
    echo "rc=${PIPESTATUS[0]}" >> gpurun_out/r2m_sanitizer.log
  done
done
cat gpurun_out/r2m_sanitizer.log | tail -60
