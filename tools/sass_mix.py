"""Opcode mix of one kernel from `ncu -i rep --page source --csv --kernel-name regex:NAME` output (first matching launch)."""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = None; ops = collections.Counter(); samp = collections.Counter(); thr = collections.Counter(); tot = tots = 0; nk = 0
for r in rows:
    if r and r[0] == "Kernel Name":
        nk += 1
        if nk > 1: break
        continue
    if r and r[0] == "Address":
        hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr): continue
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
    if not m: continue
    op = m.group(2).split(".")[0]
    try:
        n = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]]); t = int(r[ci["Thread Instructions Executed"]])
    except ValueError:
        continue
    ops[op] += n; samp[op] += s; thr[op] += t; tot += n; tots += s
print("total warp inst %.3e  samples %d" % (tot, tots))
for op, n in ops.most_common(int(sys.argv[2]) if len(sys.argv) > 2 else 25):
    print("%-10s inst %6.2f%%  samples %6.2f%%  avg lanes %.1f" % (op, 100 * n / tot, 100 * samp[op] / max(tots, 1), thr[op] / max(n, 1)))
