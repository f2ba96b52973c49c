#!/bin/bash
# round 2, call 22: bump mapping on the GPU, full suite
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== bumpmap first"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "bumpmap" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
