mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8
echo "== experiments (1 M soup, 16 spp)"
python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16"
PB2_POOL=8388608 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/pool8M /'
PB2_POOL=16777216 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/pool16M /'
PB2_POOL=2097152 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/pool2M /'
PB2_SYNC_EVERY=16 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/sync16 /'
PB2_SYNC_EVERY=4 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/sync4 /'
PB2_FINISH=0 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/finish0 /'
PB2_FINISH=1024 python tools/probe.py soup 1000000 16 "0" 2>&1 | grep "probe soup 1000000 16" | sed 's/^/finish1024 /'
echo "== ncu launch list of the bench command"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -c 200 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/r02_bench_launches.csv
python tools/sum_launches.py gpurun_out/r02_bench_launches.csv 2>&1 | tail -25
gzip -f gpurun_out/r02_bench_launches.csv
echo "== ncu full: trace (default kernel) + shade + light + gen on the 1 M soup"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w|k_wf_advance|k_wf_gen" --launch-skip 12 --launch-count 4 -o gpurun_out/r02_round_kernels -f python tools/probe.py soup 1000000 4 "0" 1 > gpurun_out/ncu_round.log 2>&1; tail -1 gpurun_out/ncu_round.log
echo "== ncu full: trace on the 10 M soup, a later (incoherent) round"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w" --launch-skip 6 --launch-count 1 -o gpurun_out/r02_trace_10m -f python tools/probe.py soup 10000000 4 "0" 1 > gpurun_out/ncu_10m.log 2>&1; tail -1 gpurun_out/ncu_10m.log
