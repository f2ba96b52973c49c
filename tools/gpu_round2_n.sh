#!/bin/bash
# round 2, call 16: image textures / alpha masks on the GPU
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
echo "== textures first"
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "textured or texture or alpha" 2>&1 | tail -25
echo "== all gpu tests"
timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -15
echo "== smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
echo "== bench c2 short (regression check of the untextured path)"
timeout 600 python bench.py --steps 3 --warmup 3 2>&1 | tail -2 | cut -c1-600
