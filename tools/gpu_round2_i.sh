mkdir -p gpurun_out; rm -f gpurun_out/*.ncu-rep
echo "== ncu full: the default trace kernel, shade, light, gen on the 1 M soup (steady-state round)"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_wf_trace_w|k_wf_advance|k_wf_gen" --launch-skip 12 --launch-count 4 -o /tmp/r02_round -f python tools/probe.py soup 1000000 4 "0" 1 > gpurun_out/ncu_round.log 2>&1; tail -1 gpurun_out/ncu_round.log
for i in 0 1 2 3; do ncu -i /tmp/r02_round.ncu-rep --page details --csv 2>/dev/null | head -1 > /dev/null; done
python - <<'PY'
import csv, subprocess
rep = "/tmp/r02_round.ncu-rep"
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], stdout=subprocess.PIPE, text=True).stdout
keep = ['Duration', 'Executed Ipc Active', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'No Eligible', 'Avg. Active Threads Per Warp', 'Registers Per Thread',
        'Achieved Occupancy', 'Theoretical Occupancy', 'Branch Efficiency', 'Warp Cycles Per Issued Instruction', 'DRAM Throughput',
        'Compute (SM) Throughput', 'L1/TEX Cache Throughput', 'L2 Cache Throughput', 'Executed Instructions', 'Grid Size', 'Memory Throughput']
rows = list(csv.reader(det.splitlines()))
out = open("gpurun_out/r02_round_kernels_ncu_summary.txt", "w")
out.write("ncu --set full --clock-control none, 1 M soup 1920x1080x4, launches 12-15 of k_wf_trace_w / k_wf_advance / k_wf_gen (one steady-state round)\n")
last = None
for r in rows[1:]:
    if len(r) >= 15:
        key = (r[0], r[4])
        if key != last:
            out.write("launch %s: %s\n" % (r[0], r[4][:110]))
            last = key
        if r[12] in keep:
            out.write("  %-38s %14s %s\n" % (r[12], r[14], r[13]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
for vals in rows[2:]:
    out.write("raw %s\n" % vals[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "raw\n")
    for i, h in enumerate(hdr):
        if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'lts__t_sector_hit_rate.pct',
                 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'gpu__time_duration.sum'):
            out.write("  %-70s %s %s\n" % (h, vals[i], rows[1][i]))
out.close()
print(open("gpurun_out/r02_round_kernels_ncu_summary.txt").read()[:6000])
PY
echo "== ncu full: default trace kernel on the 10 M soup, a later round"
timeout 900 ncu --set full --clock-control none -k regex:"k_wf_trace_w" --launch-skip 6 --launch-count 1 -o /tmp/r02_trace_10m -f python tools/probe.py soup 10000000 4 "0" 1 > gpurun_out/ncu_10m.log 2>&1; tail -1 gpurun_out/ncu_10m.log
python tools/ncu_summary.py /tmp/r02_trace_10m.ncu-rep > gpurun_out/r02_trace_10m_ncu_summary.txt 2>&1; cat gpurun_out/r02_trace_10m_ncu_summary.txt | tail -32
echo "== config 5 generator on one GPU: 50 M triangles, 3840x2160, maxdepth 16, 16 of the 1024 spp"
timeout 1500 python bench.py --tris 50000000 --jitter 0.005 --seed 5050 --xres 3840 --yres 2160 --spp 16 --maxdepth 16 --no-cpu-baseline --steps 3 --warmup 3 2>gpurun_out/bench_c5.err | tail -1 > gpurun_out/r02_bench_c5_50m_16spp_n1.json; cut -c1-400 gpurun_out/r02_bench_c5_50m_16spp_n1.json; tail -n 3 gpurun_out/bench_c5.err
free -g | head -2
