#!/usr/bin/env python
"""Offline SIMT cost model of the HEX8 projection kernels (CPU only).

Runs the host build of the solver (tests/host/libiso_host.so) with its event trace on a replica of the bench workload and replays
the traces of 32 lanes in lock step the way a warp executes them: a loop runs as long as its slowest lane, every distinct branch
taken by some lane is issued once.  Costs are estimated FP64-pipe instruction counts per solver part of the HexBox variant (COST
below: evaluation of g + Newton step 25, full evaluation 45, tangent3 95, tangent2 55, ...); the conclusions do not hinge on their exact
values -- the model reproduces the measured lane utilisation (ncu: 16.9 of 32 lanes) and the measured ranking chunk > lane-refill.  Output:
lane utilisation (useful lane-work / 32 x issued work) of (a) the chunk kernel (one warp = 32 consecutive points of an element,
lanes synchronised at iteration granularity), (b) the same with one code path for all tangent-step variants, (c) the lane-refill
kernel (lanes at different iterations of different points), and the share of the issued work per solver part."""
import argparse, ctypes as C, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from fixtures import Grid, simp_hex8
import oracle
import prune_study

COST = {"g": 25.0, "F": 45.0, "it": 30.0, "ls": 15.0, "f": 12.0, 10: 95.0, 11: 95.0, 12: 95.0, 13: 55.0, 14: 55.0, 15: 55.0, 16: 5.0, 17: 30.0}


def parse(trace):
    """trace of one pair -> (phase-1 eval count, [iteration]) with iteration = ([variant code per pass], [restore evals per line-search trial])"""
    p1 = 0; its = []; cur = None; i = 0
    while i < len(trace):
        c = trace[i]
        if c == 3:
            cur = ([], []); its.append(cur)
        elif c == 1:
            if cur is None: p1 += 1
            elif cur[1]: cur[1][-1] += 1
            else: p1 += 0          # eval_g inside eval_full is not traced separately for boxes
        elif c == 2:
            cur[1].append(0)
        elif c >= 10:
            cur[0].append(int(c))
        i += 1
    return p1, its


def iter_cost_lane(it):
    return COST["F"] + COST["it"] + sum(COST[v] for v in it[0]) + sum(COST["ls"] + n * COST["g"] + COST["f"] for n in it[1])


def iter_cost_warp(its, unified):
    """its: the iterations the active lanes execute together in lock step"""
    c = COST["F"] + COST["it"]; parts = {"F": COST["F"] + COST["it"], "T": 0.0, "R": 0.0}
    for p in range(max(len(i[0]) for i in its)):
        codes = {i[0][p] for i in its if len(i[0]) > p}
        t = (max(COST[v] for v in codes) + 30.0) if unified else sum(COST[v] for v in codes)      # + the selects of the index rotation
        c += t; parts["T"] += t
    for q in range(max(len(i[1]) for i in its)):
        ns = [i[1][q] for i in its if len(i[1]) > q]
        r = COST["ls"] + max(ns) * COST["g"] + COST["f"]
        c += r; parts["R"] += r
    return c, parts


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=24)
    ap.add_argument("--period", type=float, default=64.0)
    args = ap.parse_args()
    L = prune_study.host_lib()
    L.iso_host_trace.argtypes = [np.ctypeslib.ndpointer(np.int32, flags="C"), C.c_long]
    L.iso_host_trace_len.restype = C.c_long
    n = args.n
    X, IEN, rho = simp_hex8(n, period_frac=args.period / n)
    g = Grid(X.min(0), X.max(0), 2 * n, 3)
    rn = oracle.nodal_densities(X, IEN, rho)
    rho_t, delta = 0.5, 1.1 * g.cell_size
    re_all = rn[IEN - 1]
    crossing = np.nonzero((re_all.min(1) < rho_t) & (re_all.max(1) > rho_t))[0]
    pc = [g.AABB_min[d] + g.cell_size * np.arange(g.N[d] + 1) for d in range(3)]
    buf = np.zeros(1 << 22, dtype=np.int32)
    elems = []                                  # per element: list of (p1, iterations) per pair, in the kernel's point order
    for e in crossing:
        Xe = X[IEN[e] - 1]; re = re_all[e]; lo, hi = Xe.min(0), Xe.max(0); rng = []
        for d in range(3):
            I0 = int(np.floor(g.N[d] * ((lo[d] - delta) - g.AABB_min[d]) / (g.AABB_max[d] - g.AABB_min[d])))
            I1 = int(np.floor(g.N[d] * ((hi[d] + delta) - g.AABB_min[d]) / (g.AABB_max[d] - g.AABB_min[d])))
            rng.append(np.arange(max(I0, 0), min(I1, g.N[d]) + 1))
        K, J, I = np.meshgrid(rng[2], rng[1], rng[0], indexing="ij")
        P = np.stack([pc[0][I.ravel()], pc[1][J.ravel()], pc[2][K.ravel()]], axis=1)
        dist = np.zeros(len(P)); its = np.zeros(len(P), dtype=np.int32)
        L.iso_host_trace(buf, len(buf))
        L.iso_host_project_many(np.ascontiguousarray(Xe), np.ascontiguousarray(re), len(P), np.ascontiguousarray(P), rho_t, 1, dist, its)
        tr = buf[:L.iso_host_trace_len()]
        ends = np.nonzero(tr == 0)[0]; a = 0; pairs = []
        for b in ends:
            pairs.append(parse(tr[a:b])); a = b + 1
        assert len(pairs) == len(P)
        elems.append(pairs)
    L.iso_host_trace(buf, 0)
    npairs = sum(len(p) for p in elems)
    # ---- (a)/(b) chunk kernel
    for unified in (False, True):
        useful = issued = 0.0; parts = {"P1": 0.0, "F": 0.0, "T": 0.0, "R": 0.0}
        for pairs in elems:
            for c0 in range(0, len(pairs), 32):
                ch = pairs[c0:c0 + 32]
                w = max(p[0] for p in ch) * COST["g"]; parts["P1"] += w
                useful += sum(p[0] * COST["g"] + sum(iter_cost_lane(i) for i in p[1]) for p in ch)
                for t in range(max(len(p[1]) for p in ch)):
                    c, pr = iter_cost_warp([p[1][t] for p in ch if len(p[1]) > t], unified)
                    w += c
                    for k in pr: parts[k] += pr[k]
                issued += 32 * w
        tot = sum(parts.values())
        print("chunk kernel%s: lane utilisation %.1f%%; %.0f warp instructions issued per 32 pairs (phase 1 %.0f%%, eval_full+bookkeeping %.0f%%, tangent steps %.0f%%, line search/restoration %.0f%%)" % (
            " with ONE tangent code path" if unified else "", 100 * useful / issued, issued / npairs, *(100 * parts[k] / tot for k in ("P1", "F", "T", "R"))))
    # ---- (c) lane refill per element (phase 1 once per element), lanes mix iterations of different points
    for unified in (False, True):
        useful = issued = 0.0
        for pairs in elems:
            p1 = max(p[0] for p in pairs) * COST["g"]; issued += 32 * p1; useful += 32 * p1      # uniform, once per warp
            lanes = [None] * 32; q = 0
            while True:
                for l in range(32):
                    if lanes[l] is None and q < len(pairs):
                        lanes[l] = list(pairs[q][1]); q += 1
                        if not lanes[l]: lanes[l] = None
                act = [l for l in range(32) if lanes[l]]
                if not act: break
                c, _ = iter_cost_warp([lanes[l][0] for l in act], unified)
                issued += 32 * (c + 20.0)                                # refill bookkeeping per step
                useful += sum(iter_cost_lane(lanes[l][0]) for l in act)
                for l in act:
                    lanes[l].pop(0)
                    if not lanes[l]: lanes[l] = None
        print("lane-refill kernel%s: lane utilisation %.1f%%; %.0f warp instructions issued per 32 pairs" % (" with ONE tangent code path" if unified else "", 100 * useful / issued, issued / npairs))
    # ---- (d) persistent warps: a warp walks through many elements and refills idle lanes across element boundaries (no drain per element)
    for unified in (False, True):
        useful = issued = 0.0
        nwarps = 16
        for w in range(nwarps):
            stream = [p[1] for pairs in elems[w::nwarps] for p in pairs]
            lanes = [None] * 32; q = 0
            while True:
                for l in range(32):
                    while lanes[l] is None and q < len(stream):
                        lanes[l] = list(stream[q]) or None; q += 1
                act = [l for l in range(32) if lanes[l]]
                if not act: break
                c, _ = iter_cost_warp([lanes[l][0] for l in act], unified)
                issued += 32 * (c + 30.0)                                # refill bookkeeping + per-lane element constants
                useful += sum(iter_cost_lane(lanes[l][0]) for l in act)
                for l in act:
                    lanes[l].pop(0)
                    if not lanes[l]: lanes[l] = None
        print("persistent lane-refill%s: lane utilisation %.1f%%; %.0f warp instructions issued per 32 pairs" % (" with ONE tangent code path" if unified else "", 100 * useful / issued, issued / npairs))


if __name__ == "__main__":
    main()
