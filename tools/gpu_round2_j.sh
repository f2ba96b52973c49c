mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== ray-pool kernel: record parity on one scene first (bounded)"; timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "wavefront_trace_kernels_write and soup" 2>&1 | tail -5
echo "== all gpu tests (two pipelines are the default now)"; timeout 1500 python -m pytest tests -m gpu -q --timeout 300 2>&1 | tail -8
echo "== one vs two wavefront pipelines; ray pool"
for w in "soup 1000000 16" "file tests/scenes/killeroo_like.pbrt 16" "instanced 100000 8"; do
  PB2_PIPES=1 timeout 300 python tools/probe.py $w "0" 2>&1 | grep "spp flags" | sed 's/^/pipes1 /'
  PB2_PIPES=2 timeout 300 python tools/probe.py $w "0" 2>&1 | grep "spp flags" | sed 's/^/pipes2 /'
done
PB2_PIPES=1 timeout 300 python tools/probe.py soup 1000000 16 "128" 2>&1 | grep "spp flags" | sed 's/^/pipes1 /'
PB2_PIPES=2 timeout 300 python tools/probe.py soup 1000000 16 "128" 2>&1 | grep "spp flags" | sed 's/^/pipes2 /'
PB2_PIPES=2 timeout 300 python tools/probe.py soup 10000000 4 "0 128" 2>&1 | grep "spp flags" | sed 's/^/pipes2 /'
echo "== bench C2"; timeout 900 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/r02_bench_c2_n1_c.json; cut -c1-200 gpurun_out/r02_bench_c2_n1_c.json
