"""Per-region instruction statistics of a kernel from `ncu --page source --csv` output.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_regions.py src.csv [block]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
blk = int(sys.argv[2]) if len(sys.argv) > 2 else 20
hdr = rows[1]
ci = {h: i for i, h in enumerate(hdr)}
src, ie, te, ss = ci['Source'], ci['Instructions Executed'], ci['Thread Instructions Executed'], ci['# Samples']
body = rows[2:]
f = lambda x: float(x.replace(',', '') or 0)
tot = sum(f(r[ie]) for r in body)
tots = sum(f(r[ss]) for r in body)
print("kernel:", rows[0][1])
print("total warp instr %.0f, thread instr %.0f, avg thr %.2f" % (tot, sum(f(r[te]) for r in body), sum(f(r[te]) for r in body) / tot))
for s in range(0, len(body), blk):
    seg = body[s:s + blk]
    w = sum(f(r[ie]) for r in seg)
    t = sum(f(r[te]) for r in seg)
    smp = sum(f(r[ss]) for r in seg)
    ops = {}
    for r in seg:
        p = r[src].split()
        if not p:
            continue
        op = p[1] if p[0].startswith('@') else p[0]
        op = op.split('.')[0]
        ops[op] = ops.get(op, 0) + 1
    top = ' '.join("%s:%d" % kv for kv in sorted(ops.items(), key=lambda kv: -kv[1])[:6])
    print("%4d  instr %5.1f%%  samples %5.1f%%  thr %5.1f  %s" % (s, 100 * w / tot, 100 * smp / max(tots, 1), t / max(w, 1), top))
