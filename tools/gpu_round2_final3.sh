#!/bin/bash
# round 2: configs 3 and 4 at N = 1 with the final code (digit tables, per-scene pipeline count)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for w in killeroo instanced; do
  echo "== bench $w N=1"
  timeout 600 python bench.py --workload $w --steps 3 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_${w}_final.err | tail -1 > gpurun_out/r02_bench_${w}_n1_final.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${w}_n1_final.json'))
print('$w N=1 value %.1f e2e %.1f ms/step %.1f launches %d frac %.3f' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['gpu_launches'], d['roofline']['frac']))
PY
done
