mkdir -p gpurun_out
nvidia-smi -L
echo "== new single-GPU tests"; timeout 900 python -m pytest tests -m gpu -q -k "lazy or emissive or standalone or maxdepth" 2>&1 | tail -8
echo "== two-process test"; timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -15
run() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n "$@"; }
echo "== bench C2 N=2"; timeout 900 bash -c "$(declare -f run); run 2 --steps 3 --warmup 3" 2>gpurun_out/bench_c2_n2.err | tail -1 > gpurun_out/r02_bench_c2_n2.json; cut -c1-300 gpurun_out/r02_bench_c2_n2.json
echo "== bench C3 N=2"; timeout 900 bash -c "$(declare -f run); run 2 --steps 3 --warmup 3 --workload killeroo" 2>gpurun_out/bench_c3_n2.err | tail -1 > gpurun_out/r02_bench_c3_n2.json; cut -c1-300 gpurun_out/r02_bench_c3_n2.json
tail -n 5 gpurun_out/bench_c2_n2.err gpurun_out/bench_c3_n2.err
