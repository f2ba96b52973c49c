"""Prints the headline numbers of an .ncu-rep (details page + a few raw metrics + stall reasons)."""
import csv
import subprocess
import sys

rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], stdout=subprocess.PIPE, text=True).stdout
keep = ['Duration', 'Executed Ipc Active', 'L1/TEX Hit Rate', 'L2 Hit Rate', 'No Eligible', 'Avg. Active Threads Per Warp',
        'Registers Per Thread', 'Achieved Occupancy', 'Theoretical Occupancy', 'Branch Efficiency', 'Eligible Warps Per Scheduler',
        'Warp Cycles Per Issued Instruction', 'DRAM Throughput', 'Compute (SM) Throughput', 'L1/TEX Cache Throughput',
        'L2 Cache Throughput', 'Executed Instructions', 'Grid Size', 'Block Size', 'Memory Throughput', 'SM Frequency', 'DRAM Frequency']
rows = list(csv.reader(det.splitlines()))
name = None
for r in rows[1:]:
    if len(r) >= 15:
        if name is None:
            name = r[4]
            print("kernel:", name)
        if r[12] in keep:
            print("  %-38s %14s %s" % (r[12], r[14], r[13]))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, vals = rows[0], rows[-1]
stalls = []
for i, h in enumerate(hdr):
    if 'smsp__average_warps_issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        stalls.append((float(vals[i].replace(',', '')), h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
    if h in ('dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum',
             'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum', 'smsp__inst_executed.sum', 'smsp__thread_inst_executed.sum', 'gpu__time_duration.sum'):
        print("  %-50s %s %s" % (h, vals[i], rows[1][i]))
for v, h in sorted(stalls, reverse=True)[:7]:
    print("  stall %-28s %.2f warps/issue" % (h, v))
