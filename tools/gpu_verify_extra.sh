#!/bin/bash
# Developer script: second verification call - the whole GPU suite again after a merge, memcheck over every scene class,
# and the cost of the filtered film write on the bench workload.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== filtered film"; for f in box gaussian sinc; do PB2_SOUP_FILTER=$f PROBE_SPP=16 PROBE_TAG=filter_$f python tools/perf_probe.py 2>&1 | grep probe | tail -1; done
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitizer.log 2>&1; echo "rc $?"; grep " ok \|ERROR SUMMARY" gpurun_out/sanitizer.log | tail -25
