mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== probes"
python tools/probe.py soup 1000000 16 "0 32 4" 2>&1 | grep probe
python tools/probe.py instanced 100000 8 "0 32 4" 2>&1 | grep probe
python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0 32 4" 2>&1 | grep probe
echo "== bench C2"; timeout 900 python bench.py 2>gpurun_out/bench_c2.err | tail -1 > gpurun_out/r02_bench_c2_n1.json; cat gpurun_out/r02_bench_c2_n1.json | cut -c1-600
echo "== bench C3 killeroo"; timeout 900 python bench.py --workload killeroo 2>gpurun_out/bench_c3.err | tail -1 > gpurun_out/r02_bench_c3_n1.json; cat gpurun_out/r02_bench_c3_n1.json | cut -c1-400
echo "== bench C4 instanced"; timeout 900 python bench.py --workload instanced 2>gpurun_out/bench_c4.err | tail -1 > gpurun_out/r02_bench_c4_n1.json; cat gpurun_out/r02_bench_c4_n1.json | cut -c1-400
echo "== bench soup 10M"; timeout 1200 python bench.py --tris 10000000 2>gpurun_out/bench_10m.err | tail -1 > gpurun_out/r02_bench_soup10m_n1.json; cat gpurun_out/r02_bench_soup10m_n1.json | cut -c1-400
tail -3 gpurun_out/*.err
