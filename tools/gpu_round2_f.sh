mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== probes"
python tools/probe.py soup 1000000 16 "0 64 4" 2>&1 | grep "probe"
python tools/probe.py instanced 100000 8 "0" 2>&1 | grep "probe inst"
python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0" 2>&1 | grep "probe file"
python tools/probe.py soup 10000000 4 "0 64" 2>&1 | grep "probe"
echo "== bench C2"; timeout 900 python bench.py --no-cpu-baseline 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/r02_bench_c2_n1_b.json; cut -c1-200 gpurun_out/r02_bench_c2_n1_b.json
