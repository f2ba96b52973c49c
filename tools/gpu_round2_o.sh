#!/bin/bash
# round 2, call 17: SobolSampler on the GPU, full suite, regression bench, compute-sanitizer
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== sobol first"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sobol" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
echo "== bench c2 short (regression check)"
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | cut -c1-300
echo "== compute-sanitizer memcheck over every scene class and kernel selection (small sizes)"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
tail -4 gpurun_out/r02_sanitizer_memcheck.txt
