#!/bin/bash
# round 2, last verification on one GPU: tests, smoke, a short bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (own arm, 3 steps, no cpu baseline)"
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-330
