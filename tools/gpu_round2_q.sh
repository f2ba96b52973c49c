#!/bin/bash
# round 2, call 19: device-side share of the multi-GPU efficiency, emulated on one GPU (1/8 of the tiles)
cd "$(dirname "$0")/.."
for pool in 4194304 2097152 1048576 524288; do
  echo "== PB2_POOL=$pool"
  PB2_POOL=$pool timeout 300 python tools/probe_partition.py 1000000 64 "1 2 4 8" 4 2>&1 | grep partition
done
echo "== PB2_POOL=1048576 PB2_PIPES=1 / 4"
PB2_POOL=1048576 PB2_PIPES=1 timeout 300 python tools/probe_partition.py 1000000 64 "1 8" 4 2>&1 | grep partition
PB2_POOL=1048576 PB2_PIPES=4 timeout 300 python tools/probe_partition.py 1000000 64 "1 8" 4 2>&1 | grep partition
