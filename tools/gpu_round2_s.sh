#!/bin/bash
# round 2, call 21: environment-mapped infinite light on the GPU, full suite, regression bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== envmap first"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "envmap" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
echo "== bench c2 (regression check: the shade kernel is at 128 registers now)"
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | cut -c1-300
