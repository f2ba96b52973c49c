mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== ray pool variants with more resident warps (pipes 1)"
for v in 0 6 7 8 9; do PB2_PIPES=1 PB2_POOLVAR=$v timeout 300 python tools/probe.py soup 1000000 16 "128" 2>&1 | grep "spp flags" | sed "s/^/poolvar$v /"; done
PB2_PIPES=1 timeout 300 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "spp flags" | sed "s/^/ref /"
echo "== record parity of the pool variants"
for v in 6 9; do PB2_POOLVAR=$v timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "wavefront_trace_kernels_write and (soup or materials)" 2>&1 | tail -1; done
echo "== adaptive number of pipelines (first frame calibrates; probe renders 3 frames, best is reported)"
PB2_VERBOSE=1 timeout 300 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "spp flags\|pipelines"
PB2_VERBOSE=1 timeout 300 python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0" 2>&1 | grep "spp flags\|pipelines"
PB2_VERBOSE=1 timeout 300 python tools/probe.py instanced 100000 8 "0" 2>&1 | grep "spp flags\|pipelines"
echo "== ncu pool variant 6"
PB2_PIPES=1 PB2_POOLVAR=6 timeout 600 ncu --set full --clock-control none -k regex:"k_wf_trace_pool" --launch-skip 3 --launch-count 1 -o /tmp/pool6 -f python tools/probe.py soup 1000000 4 "128" 1 > gpurun_out/ncu_pool6.log 2>&1; tail -1 gpurun_out/ncu_pool6.log
python tools/ncu_summary.py /tmp/pool6.ncu-rep > gpurun_out/r02_pool6_ncu_summary.txt 2>&1; grep -v Frequency gpurun_out/r02_pool6_ncu_summary.txt | head -40
