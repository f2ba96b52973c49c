#!/bin/bash
# round 2: kernel parameters as __grid_constant__ (no per-thread local copy of DScene / DRenderParams): parity + bench
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -5
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench (own arm, 4 steps, no cpu baseline)"
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_gridconst_n1.json; cut -c1-330 gpurun_out/r02_bench_gridconst_n1.json
