#!/bin/bash
# round 2, call 18: the full HLBVH build on the device
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== hlbvh tests"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "hlbvh" 2>&1 | tail -15
echo "== timing 1 M"
timeout 600 python tools/hlbvh_timing.py 1000000 2>&1 | grep -v "^pb2: trace\|Warning" | tail -12
echo "== timing 10 M"
timeout 900 python tools/hlbvh_timing.py 10000000 2>&1 | grep -v "^pb2: trace\|Warning" | tail -12
