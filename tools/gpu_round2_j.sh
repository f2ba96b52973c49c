mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== all gpu tests (two pipelines are the default now)"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== one vs two wavefront pipelines"
for w in "soup 1000000 16" "file tests/scenes/killeroo_like.pbrt 16" "instanced 100000 8"; do
  PB2_PIPES=1 python tools/probe.py $w "0" 2>&1 | grep "spp flags" | sed 's/^/pipes1 /'
  PB2_PIPES=2 python tools/probe.py $w "0" 2>&1 | grep "spp flags" | sed 's/^/pipes2 /'
done
PB2_PIPES=2 PB2_POOL=8388608 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "spp flags" | sed 's/^/pipes2 pool8M /'
echo "== bench C2"; timeout 900 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/r02_bench_c2_n1_c.json; cut -c1-200 gpurun_out/r02_bench_c2_n1_c.json
