#!/bin/bash
for w in 4 8 16; do BCI_REC_WPT=$w timeout 200 python scripts/time_fp32.py 256 2>&1 | grep "train step" | sed "s/^/WPT=$w /"; done
timeout 200 python scripts/time_fp32.py 256 2>&1 | grep "train step" | sed "s/^/default /"
for w in 4 8; do BCI_REC_WPT=$w timeout 200 python scripts/time_fp32.py 128 2>&1 | grep "train step" | sed "s/^/WPT=$w /"; done
