#!/bin/bash
for B in 16896 4100; do
  timeout 200 python scripts/time_paths.py $B 2>&1 | tail -3 | sed "s/^/stream B=$B: /"
  BCI_BF16_POOL=two timeout 200 python scripts/time_paths.py $B 2>&1 | tail -3 | sed "s/^/two    B=$B: /"
done
