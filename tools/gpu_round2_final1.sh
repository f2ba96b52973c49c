#!/bin/bash
# round 2, final verification on one GPU: tests, smoke, the bench lines of both arms
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -6
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "== bench (own arm, default flags)"
timeout 900 python bench.py > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/bench_final.err; tail -1 gpurun_out/r02_bench_final_n1.json | cut -c1-1500
echo "== bench --impl reference (short)"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>/dev/null | tail -1 | cut -c1-600
