#!/bin/bash
mkdir -p gpurun_out
which compute-sanitizer; compute-sanitizer --version 2>&1 | head -3
timeout 240 compute-sanitizer --tool synccheck python scripts/sanitize.py h256 > gpurun_out/r2m0.log 2>&1; echo "rc=$?"; head -30 gpurun_out/r2m0.log
