mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== ray pool variants (pipes 1)"
for v in 0 1 2 3 4 5; do PB2_PIPES=1 PB2_POOLVAR=$v timeout 300 python tools/probe.py soup 1000000 16 "128" 2>&1 | grep "spp flags" | sed "s/^/poolvar$v /"; done
PB2_PIPES=1 timeout 300 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "spp flags" | sed "s/^/ref /"
echo "== ncu: the pool kernel"
PB2_PIPES=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_pool" --launch-skip 3 --launch-count 1 -o /tmp/pool -f python tools/probe.py soup 1000000 4 "128" 1 > gpurun_out/ncu_pool.log 2>&1; tail -1 gpurun_out/ncu_pool.log
python tools/ncu_summary.py /tmp/pool.ncu-rep > gpurun_out/r02_pool_ncu_summary.txt 2>&1; cat gpurun_out/r02_pool_ncu_summary.txt
ncu -i /tmp/pool.ncu-rep --page source --csv > /tmp/pool_src.csv 2>/dev/null; python tools/ncu_regions.py /tmp/pool_src.csv 25 > gpurun_out/r02_pool_ncu_regions.txt 2>&1; head -60 gpurun_out/r02_pool_ncu_regions.txt
echo "== number of pipelines"
for p in 1 2 3 4; do PB2_PIPES=$p timeout 300 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "spp flags" | sed "s/^/pipes$p /"; done
for p in 2 3 4; do PB2_PIPES=$p timeout 300 python tools/probe.py file tests/scenes/killeroo_like.pbrt 16 "0" 2>&1 | grep "spp flags" | sed "s/^/pipes$p /"; done
PB2_PIPES=4 PB2_POOL=8388608 timeout 300 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "spp flags" | sed "s/^/pipes4 pool8M /"
