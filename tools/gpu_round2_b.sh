mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== probes"
python tools/probe.py soup 1000000 16 "0 32 4 36" 2>&1 | grep probe
PB2_FINISH=0 python tools/probe.py soup 1000000 16 "0 4" 2>&1 | grep probe | sed 's/^/nofinish /'
python tools/probe.py soup 10000000 4 "0 32 4 36" 2>&1 | grep probe
python tools/probe.py instanced 100000 8 "0 4" 2>&1 | grep probe
echo "== ncu 1M wide4 ld256"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_w4_ld256_1m -f python tools/probe.py soup 1000000 4 "32" 1 > gpurun_out/ncu_w4_ld256_1m.log 2>&1; tail -1 gpurun_out/ncu_w4_ld256_1m.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w" --launch-skip 3 --launch-count 1 -o gpurun_out/r02_w2_ld256_1m -f python tools/probe.py soup 1000000 4 "36" 1 > gpurun_out/ncu_w2_ld256_1m.log 2>&1; tail -1 gpurun_out/ncu_w2_ld256_1m.log
