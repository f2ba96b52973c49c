#!/bin/bash
# round 2, call 20: frame-tail parameters at a 1/8 share of the frame
cd "$(dirname "$0")/.."
for fin in 64 256 1024 4096 16384; do
  echo "== PB2_FINISH=$fin"
  PB2_FINISH=$fin timeout 300 python tools/probe_partition.py 1000000 64 "1 8" 4 2>&1 | grep partition
done
for se in 2 4 16; do
  echo "== PB2_SYNC_EVERY=$se"
  PB2_SYNC_EVERY=$se timeout 300 python tools/probe_partition.py 1000000 64 "1 8" 4 2>&1 | grep partition
done
