# usage: bash tools/gpu_round2_multi.sh <N> "<workloads>" [tests]
N=$1; W="$2"; mkdir -p gpurun_out
nvidia-smi -L | wc -l
if [ -n "$3" ]; then echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -6; fi
for w in $W; do
  echo "== bench $w N=$N"
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 3 --warmup 3 --workload $w 2>gpurun_out/bench_${w}_n$N.err | tail -1 > gpurun_out/r02_bench_${w}_n$N.json
  python - <<PY
import json
d=json.load(open('gpurun_out/r02_bench_${w}_n$N.json'))
print('$w N=$N value %.1f e2e %.1f ms/step %.1f trace_ms %.1f launches %d' % (d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['trace_ms_per_frame'], d['gpu_launches']))
PY
done
tail -n 3 gpurun_out/bench_*_n$N.err | grep -v "^$\|OMP_NUM\|\*\*\*\*" | tail -5
