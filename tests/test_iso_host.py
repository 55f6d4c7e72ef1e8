"""The projection solver's device functions (rho2sdf.jl_b200/csrc/r2s_iso.cuh) compiled for the HOST and checked against the CPU
oracle: per-lane arithmetic of the CUDA kernels without a GPU.  Covers the general trilinear element (HexTri), the axis-aligned box
element (HexBox), the FAST restoration and the per-element phase 1 (the opt-in kernel variants)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SG = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)


@pytest.fixture(scope="module")
def host():
    src = os.path.join(HERE, "host", "iso_host.cpp")
    so = os.path.join(HERE, "host", "libiso_host.so")
    hdr = os.path.join(ROOT, "rho2sdf.jl_b200", "csrc", "r2s_iso.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-I/usr/local/cuda/include", src, "-o", so])
    L = C.CDLL(so)
    dp = np.ctypeslib.ndpointer(np.float64, flags="C")
    L.iso_host_project_many.argtypes = [dp, dp, C.c_long, dp, C.c_double, C.c_int, dp, np.ctypeslib.ndpointer(np.int32, flags="C")]
    L.iso_host_is_box.argtypes = [dp, dp]
    L.iso_host_project_tet4.argtypes = [dp, dp, dp, C.c_double, dp]
    return L


def many(L, Xe, re, P, rho_t, variant):
    d = np.zeros(len(P)); it = np.zeros(len(P), dtype=np.int32)
    rc = L.iso_host_project_many(np.ascontiguousarray(Xe), np.ascontiguousarray(re), len(P), np.ascontiguousarray(P), rho_t, variant, d, it)
    return rc, d, it


def random_cases(rng, n, box):
    for t in range(n):
        c = rng.uniform(-5, 5, 3); h = rng.uniform(0.1, 1.0, 3)
        if t % 2 == 0:
            h[:] = h[0]
        Xe = c + SG * h
        if not box:
            Xe = Xe + rng.uniform(-0.15, 0.15, (8, 3)) * h
        re = rng.uniform(0, 1, 8)
        if t % 3 == 0:
            re = np.clip(rng.normal(0.5, 0.4, 8), 0.001, 1)
        if t % 7 == 0:
            re = 0.5 + rng.uniform(-1, 1, 8) * 1e-3          # nearly flat field: tiny gradients
        if re.min() < 0.5 < re.max():
            yield Xe, re, c + rng.uniform(-2.2, 2.2, (32, 3)) * h, float(h.min())


def oracle_distance(x, Xe, re, rho_t=0.5):
    ok, xi, _ = oracle.project_iso_hex8(x, rho_t, Xe, re)
    return ok, np.linalg.norm(x - (np.prod(1 + SG * xi, axis=1) / 8) @ Xe)


def test_box_detection(host):
    rng = np.random.default_rng(0)
    Xe = np.array([1.0, -2.0, 0.5]) + SG * np.array([0.3, 0.7, 0.2])
    assert host.iso_host_is_box(np.ascontiguousarray(Xe), np.zeros(8)) == 1
    assert host.iso_host_is_box(np.ascontiguousarray(Xe[[1, 2, 3, 0, 5, 6, 7, 4]]), np.zeros(8)) == 0        # rotated node order: general path
    assert host.iso_host_is_box(np.ascontiguousarray(Xe + rng.uniform(-1e-9, 1e-9, (8, 3))), np.zeros(8)) == 0
    assert many(host, Xe[[1, 2, 3, 0, 5, 6, 7, 4]], rng.uniform(0, 1, 8), np.zeros((1, 3)), 0.5, 1)[0] == -2


def test_general_and_box_variants_match_the_oracle(host):
    rng = np.random.default_rng(11)
    worst = {0: 0.0, 1: 0.0}; npairs = 0
    for box in (True, False):
        for Xe, re, P, h in random_cases(rng, 300, box):
            for v in ((0, 1) if box else (0,)):
                _, d, _ = many(host, Xe, re, P, 0.5, v)
                for q in range(0, len(P), 4):
                    ok, do = oracle_distance(P[q], Xe, re)
                    if ok:
                        worst[v] = max(worst[v], abs(d[q] - do) / h); npairs += 1
    assert npairs > 2000
    assert worst[0] <= 1e-10 and worst[1] <= 1e-10, worst          # parity tolerance of the GPU tests is 1e-9 h


def test_box_variant_equals_general_variant(host):
    rng = np.random.default_rng(5)
    worst = 0.0
    for Xe, re, P, h in random_cases(rng, 1500, True):
        _, d0, it0 = many(host, Xe, re, P, 0.5, 0)
        _, d1, it1 = many(host, Xe, re, P, 0.5, 1)
        worst = max(worst, float(np.abs(d0 - d1).max()) / h)
    assert worst <= 1e-12, worst


def test_fast_restoration_and_element_phase1_change_nothing(host):
    """FAST only skips an evaluation whose outcome is known (|g| <= tolg by the remainder bound); phase 1 from the element is the
    same arithmetic done once: both must reproduce the exact variants bit for bit."""
    rng = np.random.default_rng(9)
    for Xe, re, P, h in random_cases(rng, 1500, True):
        _, d1, it1 = many(host, Xe, re, P, 0.5, 1)
        for v in (2, 3):
            _, d, it = many(host, Xe, re, P, 0.5, v)
            assert np.array_equal(d, d1) and np.array_equal(it, it1)
    for Xe, re, P, h in random_cases(rng, 500, False):
        _, d0, it0 = many(host, Xe, re, P, 0.5, 0)
        _, d4, it4 = many(host, Xe, re, P, 0.5, 4)
        assert np.array_equal(d0, d4) and np.array_equal(it0, it4)


def test_single_path_tangent_step_matches(host):
    """MODE 3 (one code path for all tangent-step cases, HexBox): same mathematics, different rounding -- distances agree far inside
    the 1e-9 h parity tolerance with the exact box variant and with the oracle."""
    rng = np.random.default_rng(21)
    worst = 0.0; worst_o = 0.0
    for Xe, re, P, h in random_cases(rng, 1500, True):
        _, d1, _ = many(host, Xe, re, P, 0.5, 1)
        for v in (5,):               # 5 = MODE 3 with element phase 1: what the kernels run
            _, d, _ = many(host, Xe, re, P, 0.5, v)
            worst = max(worst, float(np.abs(d - d1).max()) / h)
            ok, do = oracle_distance(P[0], Xe, re)
            if ok:
                worst_o = max(worst_o, abs(d[0] - do) / h)
    assert worst <= 5e-11 and worst_o <= 1e-10, (worst, worst_o)
    worst = 0.0
    for Xe, re, P, h in random_cases(rng, 800, False):      # general trilinear element: variant 8 = MODE 3 with element phase 1
        _, d0, _ = many(host, Xe, re, P, 0.5, 0)
        _, d8, _ = many(host, Xe, re, P, 0.5, 8)
        worst = max(worst, float(np.abs(d8 - d0).max()) / h)
    assert worst <= 5e-11, worst


def test_coordinate_offset_sensitivity(host):
    """Fed with global coordinates (like the reference and the oracle) the solver cancels the mesh offset in X(xi) - x: up to offset / h = 1e4
    the result moves by < 1e-10 h (beyond ~1e5 the line search -- and the oracle's -- hits the round-off floor).  The kernels therefore
    subtract node 0 of the element from the nodes and the grid point first (exact in floating point): element-local coordinates."""
    rng = np.random.default_rng(5)
    worst = {0: 0.0, 1: 0.0}
    for _ in range(400):
        c = np.round(rng.uniform(-5, 5, 3) * 2 ** 20) / 2 ** 20; h = np.round(rng.uniform(0.2, 1.0, 3) * 2 ** 20) / 2 ** 20
        re = rng.uniform(0, 1, 8)
        if not (re.min() < 0.5 < re.max()):
            continue
        U = np.round(rng.uniform(-2.2, 2.2, (16, 3)) * 2 ** 10) / 2 ** 10
        Xe, P = c + SG * h, c + U * h
        _, dref, _ = many(host, Xe, re, P, 0.5, 0)
        for v, off in ((0, 1e4), (1, 1e4)):
            _, d, _ = many(host, Xe + off, re, P + off, 0.5, v)
            worst[v] = max(worst[v], float(np.abs(d - dref).max()) / float(h.min()))
    assert worst[0] <= 1e-10 and worst[1] <= 1e-10, worst


def test_tet4_projection_matches_the_oracle(host):
    """project_tet4 (closest point on the in-element iso-polygon of a linear tetrahedron) on random tets, including nodes that sit exactly
    on the iso value: same existence verdict, same distance."""
    rng = np.random.default_rng(3)
    n = 0
    for t in range(6000):
        Xe = np.ascontiguousarray(rng.uniform(-1, 1, (4, 3)) + rng.uniform(-5, 5, 3))
        if abs(np.linalg.det(Xe[1:] - Xe[0])) < 1e-3:
            continue
        re = rng.uniform(0, 1, 4)
        if t % 5 == 0:
            re[rng.integers(4)] = 0.5
        x = np.ascontiguousarray(Xe.mean(0) + rng.uniform(-2, 2, 3))
        xp = np.zeros(3)
        ok = host.iso_host_project_tet4(Xe, np.ascontiguousarray(re), x, 0.5, xp)
        oko, xpo = oracle.project_iso_tet4(x, 0.5, Xe, re)
        assert bool(ok) == oko
        if ok:
            assert abs(np.linalg.norm(x - xp) - np.linalg.norm(x - xpo)) <= 1e-12
            n += 1
    assert n > 3000


def test_bench_replica_pair_by_pair(host):
    """A 24^3 replica of the bench workload (synthetic SIMP field, grid step h_e / 2, delta = 1.1 cells): the HexBox device code against the
    oracle on every 5th (element, grid point) pair of every crossing element -- the per-pair version of the GPU parity test, which only
    sees the minimum over the pairs of a grid point."""
    from fixtures import Grid, simp_hex8
    n = 24
    X, IEN, rho = simp_hex8(n, period_frac=64.0 / n)
    g = Grid(X.min(0), X.max(0), 2 * n, 3)
    rn = oracle.nodal_densities(X, IEN, rho)
    delta = 1.1 * g.cell_size
    re_all = rn[IEN - 1]
    crossing = np.nonzero((re_all.min(1) < 0.5) & (re_all.max(1) > 0.5))[0]
    pc = [g.AABB_min[d] + g.cell_size * np.arange(g.N[d] + 1) for d in range(3)]
    worst = 0.0; npairs = 0
    for e in crossing:
        Xe = np.ascontiguousarray(X[IEN[e] - 1]); re = np.ascontiguousarray(re_all[e]); lo, hi = Xe.min(0), Xe.max(0)
        rng = []
        for d in range(3):
            I0 = int(np.floor(g.N[d] * ((lo[d] - delta) - g.AABB_min[d]) / (g.AABB_max[d] - g.AABB_min[d])))
            I1 = int(np.floor(g.N[d] * ((hi[d] + delta) - g.AABB_min[d]) / (g.AABB_max[d] - g.AABB_min[d])))
            rng.append(np.arange(max(I0, 0), min(I1, g.N[d]) + 1))
        K, J, I = np.meshgrid(rng[2], rng[1], rng[0], indexing="ij")
        P = np.ascontiguousarray(np.stack([pc[0][I.ravel()], pc[1][J.ravel()], pc[2][K.ravel()]], axis=1)[::5])
        rc, d, _ = many(host, Xe, re, P, 0.5, 1)
        assert rc == 0
        for q in range(len(P)):
            ok, do = oracle_distance(P[q], Xe, re)
            assert ok
            worst = max(worst, abs(d[q] - do) / g.cell_size); npairs += 1
    assert npairs > 15000 and worst <= 1e-10, (npairs, worst)


def test_degenerate_pairs_found_by_fuzzing(host):
    """Two (element, point) pairs found by fuzzing the host build against the oracle (tests/golden/degenerate_pairs.npz):
    0: chapadlo element 1115, two nodal densities exactly at rho_t so that g vanishes along an element edge -- after the multiplier test
       releases a bound, the released variable is re-fixed because its numerically zero step component has the wrong sign (1e-17);
       the device code stops at a non-KKT point (distance 4.913467 instead of 4.912790);
    1: symmetric nodal densities on a 0.25 lattice -- two bounds block the step at the same ratio, the tie is broken by the last bit
       and leads to another local minimum (0.8192 vs 0.7865).
    Fix (to be made in r2s_iso.cuh AND oracle/r2s_oracle.c together, then verified on the GPU): do not re-fix on a component with
    |d_i| <= tolx (set it to zero instead); break ratio-test ties within 1e-12 towards the lower index."""
    c = np.load(os.path.join(HERE, "golden", "degenerate_pairs.npz"))
    for k in range(2):
        Xe, re, x, v = np.ascontiguousarray(c["Xe"][k]), np.ascontiguousarray(c["re"][k]), c["x"][k], int(c["variant"][k])
        _, d, _ = many(host, Xe, re, x[None, :], 0.5, v)
        ok, do = oracle_distance(x, Xe, re)
        assert ok and abs(d[0] - do) <= 1e-9 * float((Xe.max(0) - Xe.min(0)).min())
