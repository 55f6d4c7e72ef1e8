"""Hottest source lines of one kernel: ncu -i rep --page source --csv --print-source cuda,sass --kernel-name regex:NAME | source_hot.py [N]"""
import csv, sys
rows = list(csv.reader(sys.stdin)); hdr = None; out = []; fname = ""; first = None; nfun = 0
for r in rows:
    if r and r[0] == "File Path": fname = r[1].split("/")[-1]; continue
    if r and r[0] == "Function Name":
        nfun = nfun + 1 if fname == first or first is None else nfun
        first = first or fname
        if nfun > 1 and fname == first: break          # first launch only (a launch lists one block per source file)
        continue
    if r and r[0] == "Line No" and "# Samples" in r:
        hdr = r; ci = {h: i for i, h in enumerate(hdr)}; continue
    if hdr is None or len(r) < len(hdr) or not r[0]: continue
    try: s = int(r[ci["# Samples"]]); n = int(r[ci["Instructions Executed"]]); t = int(r[ci["Thread Instructions Executed"]])
    except ValueError: continue
    out.append((s, n, t, fname + ":" + r[0], r[1][:140]))
tot = sum(o[0] for o in out) or 1; toti = sum(o[1] for o in out) or 1
print("samples %d  warp inst %.3e" % (tot, toti))
for s, n, t, l, src in sorted(out, reverse=True)[:int(sys.argv[1]) if len(sys.argv) > 1 else 30]:
    print("%5.1f%% smp %5.1f%% inst lanes %4.1f  L%s: %s" % (100 * s / tot, 100 * n / toti, t / max(n, 1), l, src))
