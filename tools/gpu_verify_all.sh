#!/bin/bash
# Developer script: everything that needs a GPU after a batch of CPU-side changes, in ONE gpurun call.
mkdir -p gpurun_out
echo "== pytest -m gpu"; timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for s in lights uber; do echo "== $s vs oracle"; python tools/gpu_check.py tests/scenes/$s.pbrt 2>&1 | grep "intersect:\|li:\|image:\|sample \["; done
echo "== smoke"; python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== probes"; for i in 1 2; do PROBE_SPP=16 PROBE_TAG=c2_$i python tools/perf_probe.py 2>&1 | grep probe | tail -1; done
PROBE_TAG=c4 python tools/perf_probe_instanced.py 2>&1 | tail -1
python tools/hlbvh_probe.py 1000000 10000000 2>&1 | grep "hlbvh probe"
echo "== bench"; timeout 900 python bench.py 2>/dev/null | tail -1 > gpurun_out/bench_n1.json
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['trace_share_of_step'], d['gpu_launches'], d['clocks'], d['cpu_baseline']['value'])"
echo "== filtered film"; for f in gaussian sinc; do PB2_SOUP_FILTER=$f PROBE_SPP=16 PROBE_TAG=filter_$f python tools/perf_probe.py 2>&1 | grep probe | tail -1; done
[ -n "$SKIP_NCU" ] && exit 0
echo "== ncu full"; PROBE_SPP=4 PROBE_ITERS=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_wf_advance|k_wf_trace_w" --launch-skip 4 --launch-count 3 -o gpurun_out/final_kernels -f python tools/perf_probe.py > gpurun_out/ncu_final.log 2>&1; tail -1 gpurun_out/ncu_final.log
echo "== ncu launch list of the bench command"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/bench_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -c 300 gpurun_out/bench_under_ncu.log; wc -l gpurun_out/bench_launches.csv
echo "== memcheck"; timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_small.py > gpurun_out/sanitizer.log 2>&1; echo "rc $?"; grep " ok \|ERROR SUMMARY" gpurun_out/sanitizer.log | tail -25
