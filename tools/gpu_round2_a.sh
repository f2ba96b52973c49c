mkdir -p gpurun_out
echo "== wavefront tests"; timeout 900 python -m pytest tests -m gpu -x -q -k "wavefront" 2>&1 | tail -15
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== probes"
python tools/probe.py soup 1000000 16 "0 4 2" 2>&1 | grep probe
python tools/probe.py instanced 100000 8 "0 4 2" 2>&1 | grep probe
python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0 4 2" 2>&1 | grep probe
python tools/probe.py soup 10000000 4 "0 4 2" 2>&1 | grep probe
echo "== ncu 1M wide4"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_w4_1m -f python tools/probe.py soup 1000000 4 "0" 1 > gpurun_out/ncu_w4_1m.log 2>&1; tail -2 gpurun_out/ncu_w4_1m.log
echo "== ncu 10M wide4"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_w4_10m -f python tools/probe.py soup 10000000 2 "0" 1 > gpurun_out/ncu_w4_10m.log 2>&1; tail -2 gpurun_out/ncu_w4_10m.log
