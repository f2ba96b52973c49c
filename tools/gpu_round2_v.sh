#!/bin/bash
# round 2, call 24: texture combinators + the chain flag test on the GPU, full suite, memcheck
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== texcombine / chain first"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "texcombine or chained" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
echo "== compute-sanitizer memcheck"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_small.py > gpurun_out/r02_sanitizer_memcheck.txt 2>&1; echo "memcheck rc=$?"
grep " ok \|ERROR SUMMARY" gpurun_out/r02_sanitizer_memcheck.txt | tail -8
