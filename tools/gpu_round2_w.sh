#!/bin/bash
# round 2, call 26: the chained trace kernel for the rounds after the work counter has run out (drain phase)
cd "$(dirname "$0")/.."
for dc in 0 1; do
  echo "== PB2_DRAIN_CHAIN=$dc"
  PB2_DRAIN_CHAIN=$dc timeout 300 python tools/probe_partition.py 1000000 64 "1 2 4 8" 4 2>&1 | grep partition
done
echo "== gpu tests with the drain switch (default on)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -5
