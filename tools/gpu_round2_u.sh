#!/bin/bash
# round 2, call 23: the light step inside the trace kernel (CHAIN): parity through the whole GPU suite, then A/B timing
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== gpu tests with the chain kernels (default)"
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -15
for chain in 0 1; do
  echo "== PB2_CHAIN=$chain"
  PB2_CHAIN=$chain timeout 300 python tools/probe.py soup 1000000 16 "0" 4 2>&1 | grep "probe soup.*flags"
  PB2_CHAIN=$chain timeout 300 python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0" 4 2>&1 | grep "probe file.*flags"
  PB2_CHAIN=$chain timeout 300 python tools/probe.py instanced 100000 8 "0" 4 2>&1 | grep "probe instanced.*flags"
  PB2_CHAIN=$chain timeout 300 python tools/probe_partition.py 1000000 64 "1 8" 3 2>&1 | grep partition
done
