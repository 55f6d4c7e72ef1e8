"""CPU tests (no GPU): the oracle against the reference's own golden values and analytic answers.

Goldens: test/HexBlockSdfTest.jl:25-32, test/HexSphereSdfTest.jl:26-35 (reference repo).  SURVEY.md section 4 established that
the *_sdf goldens were produced with a band half-width delta in (2,3) cells (2.5 fits) while the shipped source has 1.1
(src/SignedDistances/sdfOnDensityField.jl:158); both settings are pinned here.
"""
import numpy as np
import pytest

import oracle
from fixtures import BLOCK_RHO_N, Grid, block_geometry, load_mesh, schlafli_tet4, simp_hex8


def isapprox(a, b, rtol=0.0, atol=0.0):
    return abs(a - b) <= max(atol, rtol * max(abs(a), abs(b)))


@pytest.fixture(scope="module")
def sphere():
    X, IEN, rho = load_mesh("sphere")
    return X, IEN, rho, oracle.nodal_densities(X, IEN, rho)


def test_sphere_nodal_density_goldens(sphere):
    X, IEN, rho, rn = sphere
    assert isapprox(rn.max(), 1.0000000000000022, rtol=1e-10, atol=1e-12)        # HexSphereSdfTest.jl:26,85
    assert isapprox(rn.mean(), 0.29490556408887564, rtol=1e-10, atol=1e-12)      # :27,86
    assert rn.min() >= 0.0 and rn.max() <= 1.1 and rn.std() > 0.1                 # :160-166


def test_sphere_sdf_goldens_delta_2p5(sphere):
    X, IEN, rho, rn = sphere
    g = Grid(X.min(0), X.max(0), 10, 3)
    assert list(g.N) == [16, 16, 16] and g.ngp == 4913
    d, xp, st = oracle.eval_distances(X, IEN, g, rn, 0.5, 2.5)
    s = oracle.sign_detection(X, IEN, g, rn, 0.5)
    sdf = d * s
    assert st["not_converged"] == 0
    assert isapprox(sdf.max(), 0.8669785608800439, rtol=1e-10, atol=1e-12)       # :28,137
    assert isapprox(sdf.mean(), -3.7370242217627172e9, atol=1e5)                  # :29,140
    assert int((sdf < -1e9).sum() - (sdf > 1e9).sum()) == 1836
    assert (d >= 0).all() and set(np.unique(s)) <= {-1.0, 1.0}
    assert (s > 0).sum() < (s < 0).sum() and int((s > 0).sum()) == 365


def test_sphere_sdf_shipped_delta_1p1(sphere):
    X, IEN, rho, rn = sphere
    g = Grid(X.min(0), X.max(0), 10, 3)
    d, _, st = oracle.eval_distances(X, IEN, g, rn, 0.5, 1.1, want_xp=False)
    s = oracle.sign_detection(X, IEN, g, rn, 0.5)
    sdf = d * s
    assert int((sdf < -1e9).sum() - (sdf > 1e9).sum()) == 3205                    # SURVEY.md 8c pin
    assert isapprox(sdf.max(), 0.8679117863499797, rtol=1e-12)                    # SURVEY.md 8c pin (exact KKT point)
    re = rn[IEN - 1]
    assert int((re.min(1) >= 0.5).sum()) == 208 and int(((re.min(1) < 0.5) & (re.max(1) > 0.5)).sum()) == 368


def test_block_goldens():
    X, IEN, rho = block_geometry([2, 1, 1])
    g = Grid(X.min(0), X.max(0), 20, 3)
    assert g.ngp == 7803
    d, _, _ = oracle.eval_distances(X, IEN, g, BLOCK_RHO_N, 0.5, 2.5)
    s = oracle.sign_detection(X, IEN, g, BLOCK_RHO_N, 0.5)
    sdf = d * s
    assert isapprox(sdf.max(), 0.4242640687119285, rtol=1e-10, atol=1e-12)       # HexBlockSdfTest.jl:25 (= 0.3*sqrt(2))
    assert isapprox(sdf.mean(), -1.4699474563515213e9, atol=1e5)                  # :26
    assert int((sdf < -1e9).sum() - (sdf > 1e9).sum()) == 1147
    d11, _, _ = oracle.eval_distances(X, IEN, g, BLOCK_RHO_N, 0.5, 1.1)
    assert int(((d11 * s) < -1e9).sum() - ((d11 * s) > 1e9).sum()) == 3747


@pytest.mark.parametrize("thr", [0.1, 0.9])
def test_sphere_edge_case_thresholds(sphere, thr):                                # HexSphereSdfTest.jl:169-199
    X, IEN, rho, rn = sphere
    g = Grid(X.min(0), X.max(0), 5, 3)
    d, _, _ = oracle.eval_distances(X, IEN, g, rn, thr, 1.1)
    s = oracle.sign_detection(X, IEN, g, rn, thr)
    sdf = d * s
    assert len(sdf) == g.ngp and np.isfinite(sdf).all()


def test_iso_projection_matches_published_slsqp(sphere):
    """The reference solves the projection with NLopt :LD_SLSQP (ComputeCoordsOnIso.jl:19-79).  scipy's SLSQP is the same
    published algorithm (Kraft 1988); converged tightly from the same start it must land on the oracle's point."""
    from scipy.optimize import minimize
    X, IEN, rho, rn = sphere
    SG = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    shape = lambda xi: 0.125 * np.prod(1 + SG * xi, axis=1)

    def dshape(xi):
        t = 1 + SG * xi
        return 0.125 * np.stack([SG[:, 0] * t[:, 1] * t[:, 2], SG[:, 1] * t[:, 0] * t[:, 2], SG[:, 2] * t[:, 0] * t[:, 1]], axis=1)

    rng = np.random.default_rng(0)
    re_all = rn[IEN - 1]
    cross = np.where((re_all.min(1) < 0.5) & (re_all.max(1) > 0.5))[0]
    checked = 0
    for _ in range(60):
        e = rng.choice(cross)
        Xe, re = X[IEN[e] - 1], re_all[e]
        x = Xe.mean(0) + rng.uniform(-2, 2, 3) * 0.2
        ok, xi, _ = oracle.project_iso_hex8(x, 0.5, Xe, re)
        assert ok
        r = minimize(lambda q: np.sum((x - Xe.T @ shape(q)) ** 2), np.zeros(3), jac=lambda q: -2 * (dshape(q).T @ Xe) @ (x - Xe.T @ shape(q)),
                     method="SLSQP", bounds=[(-1, 1)] * 3, constraints=[{"type": "eq", "fun": lambda q: re @ shape(q) - 0.5, "jac": lambda q: dshape(q).T @ re}],
                     options={"ftol": 1e-15, "maxiter": 1000})
        if abs(re @ shape(r.x) - 0.5) < 1e-9:
            d1, d2 = np.linalg.norm(x - Xe.T @ shape(xi)), np.linalg.norm(x - Xe.T @ shape(r.x))
            assert abs(d1 - d2) <= 1e-7 * 0.2
            checked += 1
    assert checked >= 50


def test_threaded_reference_scheme_agrees_on_smooth_field():
    """nthreads > 1 reproduces the reference's per-thread buffers (sdfOnDensityField.jl:183-195); away from the
    order-dependent boundary-face rule the merged result equals the single-buffer one."""
    X, IEN, rho = simp_hex8(8)
    rn = oracle.nodal_densities(X, IEN, rho)
    g = Grid(X.min(0), X.max(0), 16, 3)
    d1, _, _ = oracle.eval_distances(X, IEN, g, rn, 0.5, 1.1, nthreads=1, want_xp=False)
    d4, _, _ = oracle.eval_distances(X, IEN, g, rn, 0.5, 1.1, nthreads=4, want_xp=False)
    assert (np.abs(d1 - d4) > 1e-12).mean() < 0.02


def test_artifact_removal_matches_ndimage():
    from scipy import ndimage
    rng = np.random.default_rng(3)

    class G:
        N = np.array([20, 14, 9]); ngp = 21 * 15 * 10
    f = ndimage.gaussian_filter(rng.standard_normal((10, 15, 21)), 1.2)
    sdf = (f - 0.02).ravel().copy()
    out, nf = oracle.remove_artifacts(sdf, G, 0.0, 0.05)
    lab, n = ndimage.label(f - 0.02 >= 0)
    sizes = ndimage.sum(np.ones_like(lab), lab, range(1, n + 1))
    largest = sizes.max(); ms = max(1, int(np.round(0.05 * largest)))        # np.round is half-to-even like Julia's round
    keep = np.zeros(n + 1, bool); keep[1:] = (sizes >= ms) | (np.arange(1, n + 1) == 1 + int(np.argmax(sizes)))
    flip = (lab > 0) & ~keep[lab]
    assert nf == int(flip.sum()) and nf > 0
    exp = sdf.copy(); exp[flip.ravel()] = -np.abs(exp[flip.ravel()])
    assert np.array_equal(out, exp)


@pytest.mark.parametrize("N,tol", [(16, 0.10), (32, 0.05), (64, 0.02)])
def test_volume_convergence_sphere(N, tol):                                      # test/ConvergenceTests/SphereConvergenceTest.jl:364-398
    ax = np.linspace(-1, 1, N + 1, dtype=np.float32)
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    sdf = (0.5 - np.sqrt(x * x + y * y + z * z)).astype(np.float32)
    v = oracle.volume_from_sdf(sdf, np.float32(ax[1] - ax[0]), order=9)
    assert abs(v - np.pi / 6) / (np.pi / 6) < tol


@pytest.mark.parametrize("N,tol", [(16, 0.05), (32, 0.02)])
def test_volume_convergence_cube(N, tol):                                        # test/ConvergenceTests/CubeConvergenceTest.jl:392-425
    ax = np.linspace(-1, 1, N + 1, dtype=np.float32)
    z, y, x = np.meshgrid(ax, ax, ax, indexing="ij")
    sdf = (0.5 - np.maximum(np.maximum(np.abs(x), np.abs(y)), np.abs(z))).astype(np.float32)
    v = oracle.volume_from_sdf(sdf, np.float32(ax[1] - ax[0]), order=9)
    assert abs(v - 1.0) < tol


def test_threshold_search_cantilever():
    X, IEN, rho = load_mesh("cantilever_beam_vfrac_03")
    vd, vf = oracle.mesh_volume(X, IEN, rho)
    assert abs(vd - 4800.0) < 1e-6 and abs(vf - 0.3017352557218034) < 1e-9
    rn = oracle.nodal_densities(X, IEN, rho)
    rt = oracle.find_threshold(X, IEN, rn, vd * vf)
    v = oracle.isocontour_volume(X, IEN, rn, rt)
    assert abs(v - vd * vf) / (vd * vf) < 1e-4                                    # Isocontour_volume.jl:79,128
    with pytest.raises(RuntimeError):
        oracle.find_threshold(X, IEN, rn, 2 * vd)                                 # :93-95


def test_rbf_oracle_faithful_vs_ideal_lattice():
    """The faithful restatement (kernel values from the reference's Float32 coordinates) and the ideal-lattice stencil the
    GPU uses differ only at Float32 round-off."""
    X, IEN, rho = block_geometry([2, 1, 1])
    g = Grid(X.min(0), X.max(0), 12, 3)
    d, _, _ = oracle.eval_distances(X, IEN, g, BLOCK_RHO_N, 0.5, 1.1, want_xp=False)
    sdf = d * oracle.sign_detection(X, IEN, g, BLOCK_RHO_N, 0.5)
    vd, vf = oracle.mesh_volume(X, IEN, rho)
    f0, i0 = oracle.rbf_smoothing(sdf, g, True, 2, vd * vf, mode=0)
    f1, i1 = oracle.rbf_smoothing(sdf, g, True, 2, vd * vf, mode=1)
    assert i0["cg_iters"] == i1["cg_iters"] > 5
    assert np.abs(f0 - f1).max() <= 2e-4 * g.cell_size
    assert f0.shape == tuple(int(n) * 2 + 1 for n in g.N[::-1])


def test_tet4_oracle_agrees_with_hex8_on_linear_field():
    """Schlaefli split (test/PrimitiveGeometriesTest/SimpleCubeWithSchlafli.jl:22-29) of a cube with a LINEAR nodal density:
    the iso-surface is the same plane for both element types."""
    n = 4
    m = n + 1
    k, j, i = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    X = np.stack([i.ravel(), j.ravel(), k.ravel()], 1).astype(float)
    nid = lambda a, b, c: (c * m + b) * m + a + 1
    hexes = []
    for kk in range(n):
        for jj in range(n):
            for ii in range(n):
                hexes.append([nid(ii, jj, kk), nid(ii + 1, jj, kk), nid(ii + 1, jj + 1, kk), nid(ii, jj + 1, kk),
                              nid(ii, jj, kk + 1), nid(ii + 1, jj, kk + 1), nid(ii + 1, jj + 1, kk + 1), nid(ii, jj + 1, kk + 1)])
    H = np.array(hexes, dtype=np.int64)
    sch = np.array([[1, 2, 3, 7], [1, 6, 2, 7], [1, 3, 4, 7], [1, 4, 8, 7], [1, 5, 6, 7], [1, 8, 5, 7]]) - 1
    T = np.concatenate([H[:, s] for s in sch], axis=0)
    rn = 0.1 + (X @ np.array([0.11, 0.07, 0.05]))
    g = Grid(X.min(0), X.max(0), 8, 3)
    dh, _, _ = oracle.eval_distances(X, H, g, rn, 0.5, 1.1, want_xp=False)
    dt, _, _ = oracle.eval_distances(X, T, g, rn, 0.5, 1.1, want_xp=False)
    sh = oracle.sign_detection(X, H, g, rn, 0.5)
    st = oracle.sign_detection(X, T, g, rn, 0.5)
    both = (dh < 1e9) & (dt < 1e9)
    assert both.sum() > 100
    nz, ny, nx = (int(v) + 1 for v in g.N[::-1])
    kk, jj, ii = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    P = g.AABB_min + g.cell_size * np.stack([ii.ravel(), jj.ravel(), kk.ravel()], 1)
    nrm = np.array([0.11, 0.07, 0.05])
    plane = np.abs(P @ nrm + 0.1 - 0.5) / np.linalg.norm(nrm)
    # well inside the mesh and close to the plane the nearest iso point is the foot on the plane for both element types
    interior = both & (P.min(1) > 1.1) & (P.max(1) < n - 1.1) & (plane < 0.4)
    assert interior.sum() > 5
    assert np.abs(dh[interior] - plane[interior]).max() < 1e-9 and np.abs(dt[interior] - plane[interior]).max() < 1e-9
    inside = (P.min(1) > 0.01) & (P.max(1) < n - 0.01) & (np.abs(P @ nrm + 0.1 - 0.5) > 1e-6)
    assert np.array_equal(sh[inside], st[inside])


def test_inverse_map_matches_published_lbfgs():
    """The reference finds local coordinates with NLopt :LD_LBFGS, bounds +-1.1, 9 starting points, best objective
    (FindLocalCoordinates.jl:27-37,71-104).  scipy's L-BFGS-B is the same published algorithm (Byrd-Lu-Nocedal-Zhu); from the same
    starts, converged tightly, its best point must coincide with the oracle's Newton inverse map wherever the point lies inside
    the bounds -- and where it does not, both must say 'outside' (max|xi| >= 1.01, SignDetection.jl:48)."""
    from scipy.optimize import minimize
    X, IEN, rho = load_mesh("chapadlo")                      # unstructured, non-affine hexes
    SG = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    shape = lambda xi: 0.125 * np.prod(1 + SG * xi, axis=1)

    def dshape(xi):
        t = 1 + SG * xi
        return 0.125 * np.stack([SG[:, 0] * t[:, 1] * t[:, 2], SG[:, 1] * t[:, 0] * t[:, 2], SG[:, 2] * t[:, 0] * t[:, 1]], axis=1)

    starts = [(0, 0, 0)] + [tuple(0.5 * s for s in c) for c in SG]
    rng = np.random.default_rng(7)
    inside = outside = 0
    for _ in range(120):
        e = int(rng.integers(IEN.shape[0]))
        Xe = X[IEN[e] - 1]
        size = np.ptp(Xe, axis=0)
        x = Xe.mean(0) + rng.uniform(-0.6, 0.6, 3) * size
        ok, xi = oracle.inverse_map_hex8(x, Xe)
        best = None
        for s0 in starts:
            r = minimize(lambda q: np.sum((Xe.T @ shape(q) - x) ** 2), np.array(s0, float), jac=lambda q: 2 * (dshape(q).T @ Xe) @ (Xe.T @ shape(q) - x),
                         method="L-BFGS-B", bounds=[(-1.1, 1.1)] * 3, options={"ftol": 1e-16, "gtol": 1e-14, "maxiter": 500})
            if best is None or r.fun < best.fun:
                best = r
        m_ref = np.abs(best.x).max()
        if best.fun < 1e-16 * float(size @ size) and m_ref < 1.0999:      # the bounded minimiser is an exact root strictly inside the bounds
            assert ok and np.abs(xi - best.x).max() < 1e-6
            inside += 1
        else:                                                          # clamped at the bounds: the point is outside the element for both
            assert (not ok) or np.abs(xi).max() >= 1.01
            assert m_ref >= 1.01
            outside += 1
    assert inside >= 25 and outside >= 5, (inside, outside)


def _rbf_reference_literal(sdf, g, is_interp, smooth, rbf_cut=1e-3):
    """Line-by-line numpy/scipy restatement of RBFs_smoothing up to the fine-grid evaluation (src/SdfSmoothing/RBFs4Smoothing.jl:
    process_vector :15-22, create_grid :36-46, create_smooth_grid :60-74, kernel :103-107, compute_sparse_kernel_matrix :142-176 with a
    KD-tree `inrange`, IterativeSolvers.cg defaults :199, rbf_interpolation_kdtree :219-248 with a 124-nearest-neighbour query).
    Third-party pieces as in the reference: a KD-tree (scipy.spatial.cKDTree for NearestNeighbors.jl) and a sparse Float32 matrix."""
    from scipy.spatial import cKDTree
    from scipy.sparse import coo_matrix
    f32 = np.float32
    v = sdf.astype(f32)
    fin = np.abs(v) < f32(1e9)
    maxv = np.abs(v[fin]).max()
    far = np.abs(np.abs(v) - f32(1e10)) <= np.sqrt(np.finfo(f32).eps) * np.maximum(np.abs(v), f32(1e10))     # isapprox, default rtol
    v = np.where(far, np.sign(v) * maxv, v).astype(f32)
    n = [int(t) + 1 for t in g.N]
    lo, hi = g.AABB_min.astype(f32), g.AABB_max.astype(f32)
    ax = [np.linspace(np.float64(lo[d]), np.float64(hi[d]), n[d]).astype(f32) for d in range(3)]            # range(Float32, Float32, length)
    kk, jj, ii = np.meshgrid(np.arange(n[2]), np.arange(n[1]), np.arange(n[0]), indexing="ij")
    P = np.stack([ax[0][ii.ravel()], ax[1][jj.ravel()], ax[2][kk.ravel()]], axis=1)                         # x fastest, Float32
    sigma = np.float64(g.cell_size)

    def kern(a, b):                                                                                           # :103-107
        d = a - b                                                                                             # Float32
        r = np.sqrt((d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1] + d[..., 2] * d[..., 2]).astype(f32)).astype(f32)
        val = np.exp(-(r.astype(np.float64) / sigma) ** 2)
        return np.where(val > rbf_cut, val, 0.0)

    tree = cKDTree(P.astype(np.float64))
    w = v.copy()
    iters = 0
    if is_interp:
        radius = sigma * np.sqrt(-np.log(rbf_cut))
        pairs = tree.query_ball_point(P.astype(np.float64), radius)
        I = np.concatenate([np.full(len(p), i) for i, p in enumerate(pairs)]); J = np.concatenate([np.asarray(p, dtype=np.int64) for p in pairs])
        V = kern(P[I], P[J]).astype(f32)
        keep = V > f32(rbf_cut)
        K = coo_matrix((V[keep], (I[keep], J[keep])), shape=(len(P), len(P))).tocsr().astype(f32)
        # IterativeSolvers.cg(K, b): x0 = 0, reltol = sqrt(eps(Float32)), maxiter = n  (CGIterable)
        x = np.zeros(len(P), f32); r = v.copy(); u = np.zeros(len(P), f32)
        res = f32(np.linalg.norm(r)); prev = f32(1.0); tol = f32(np.sqrt(np.finfo(f32).eps)) * res
        while iters < len(P) and not (res <= tol):
            beta = f32(res * res / (prev * prev))
            u = (r + beta * u).astype(f32)
            c = (K @ u).astype(f32)
            alpha = f32(res * res / f32(np.dot(u, c)))
            x = (x + alpha * u).astype(f32); r = (r - alpha * c).astype(f32)
            prev = res; res = f32(np.linalg.norm(r)); iters += 1
        w = x

    def evaluate(Q):                                                                                          # :219-248
        maxd = f32(np.sqrt(-np.log(rbf_cut) * sigma ** 2))
        dist, idx = tree.query(Q.astype(np.float64), k=min(124, len(P)))
        out = np.zeros(len(Q), f32)
        for t in range(dist.shape[1]):                                                                         # neighbours in ascending distance
            dt = dist[:, t].astype(f32)
            term = w[idx[:, t]].astype(np.float64) * np.exp(-(dt.astype(np.float64) / sigma) ** 2)
            out = np.where(dt <= maxd, (out.astype(np.float64) + term).astype(f32), out)
        return out

    lsf = evaluate(P)
    nf = [int(t) * smooth + 1 for t in g.N]
    dx = f32((hi[0] - lo[0]) / f32(nf[0] - 1))
    fax = [(lo[d] + np.arange(nf[d], dtype=f32) * dx).astype(f32) for d in range(3)]                         # xmin + (i-1)*dx in Float32
    kk, jj, ii = np.meshgrid(np.arange(nf[2]), np.arange(nf[1]), np.arange(nf[0]), indexing="ij")
    Q = np.stack([fax[0][ii.ravel()], fax[1][jj.ravel()], fax[2][kk.ravel()]], axis=1)
    return w, lsf, evaluate(Q).reshape(nf[2], nf[1], nf[0]), iters


@pytest.mark.parametrize("interp", [True, False])
def test_rbf_oracle_matches_literal_kdtree_restatement(interp):
    """The oracle's smoothing (stencil form) against the literal KD-tree / sparse-matrix / CG formulation of the reference."""
    X, IEN, rho = block_geometry([2, 1, 1])
    g = Grid(X.min(0), X.max(0), 10, 3)
    d, _, _ = oracle.eval_distances(X, IEN, g, BLOCK_RHO_N, 0.5, 1.1, want_xp=False)
    sdf = d * oracle.sign_detection(X, IEN, g, BLOCK_RHO_N, 0.5)
    vd, vf = oracle.mesh_volume(X, IEN, rho)
    w, lsf, fine_lsf, iters = _rbf_reference_literal(sdf, g, interp, 2)
    ofine, info = oracle.rbf_smoothing(sdf, g, interp, 2, vd * vf, mode=0, want_aux=True)
    scale = float(np.abs(lsf).max())
    if interp:
        assert abs(iters - info["cg_iters"]) <= 1
    assert np.abs(w - info["weights"]).max() <= 2e-4 * float(np.abs(w).max())
    assert np.abs(lsf - info["lsf"]).max() <= 1e-4 * scale
    assert np.abs((fine_lsf + np.float32(info["th"])) - ofine).max() <= 1e-4 * scale


def test_tet4_projection_matches_published_slsqp():
    """TET4 branch of compute_coords_on_iso (ComputeCoordsOnIso.jl:90-181): min |x - X N(l)|^2 s.t. rho.N(l) = rho_t, 0 <= l <= 1,
    sum l <= 1, NLopt :LD_SLSQP from the centroid.  The problem is a convex QP, so its minimiser is unique: scipy's SLSQP (the same
    published algorithm) must land on the point the oracle gets in closed form (projection on the polygon where the plane rho = rho_t cuts the tet)."""
    from scipy.optimize import minimize
    rng = np.random.default_rng(11)
    Nf = lambda l: np.array([l[0], l[1], l[2], 1.0 - l[0] - l[1] - l[2]])          # ShapeFunctions.jl:39-73
    dN = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -1, -1]], float)
    checked = 0
    for _ in range(80):
        Xe = rng.uniform(0, 1, (4, 3)) + np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]])      # a well-shaped random tet
        re = rng.uniform(0, 1, 4)
        if not (re.min() < 0.5 < re.max()):
            continue
        x = Xe.mean(0) + rng.uniform(-1.5, 1.5, 3)
        ok, xp = oracle.project_iso_tet4(x, 0.5, Xe, re)
        assert ok
        cons = [{"type": "eq", "fun": lambda l: re @ Nf(l) - 0.5, "jac": lambda l: dN.T @ re},
                {"type": "ineq", "fun": lambda l: 1.0 - l.sum(), "jac": lambda l: -np.ones(3)}]
        r = minimize(lambda l: np.sum((x - Xe.T @ Nf(l)) ** 2), np.full(3, 0.25), jac=lambda l: -2 * (dN.T @ Xe) @ (x - Xe.T @ Nf(l)),
                     method="SLSQP", bounds=[(0, 1)] * 3, constraints=cons, options={"ftol": 1e-16, "maxiter": 500})
        if r.success and abs(re @ Nf(r.x) - 0.5) < 1e-10:
            assert np.linalg.norm(Xe.T @ Nf(r.x) - xp) < 1e-6
            checked += 1
    assert checked >= 40


def test_tet4_sign_matches_literal_restatement():
    """Sign_Detection_TET4 (SignDetection.jl:88-165) restated literally in numpy -- cell lists of create_grid_tetrahedra_mapping_TET4
    (:168-217, single thread = ascending element order), is_point_in_tetrahedron (:220-242, 4x4 solve, tol 1e-10),
    find_local_coordinates(TET4) (FindLocalCoordinates.jl:110-149, 3x3 solve + validate_local_coords: all >= 0, sum <= 1.0),
    first containing tet with rho >= rho_t wins -- against the oracle on a Schlaefli-split cube."""
    n = 4
    X, IEN, rn = schlafli_tet4(n, "radial")
    g = Grid(X.min(0), X.max(0), 2 * n, 3)
    so = oracle.sign_detection(X, IEN, g, rn, 0.5).reshape([int(v) + 1 for v in g.N[::-1]])
    dims = g.N + 1
    tets = X[IEN - 1]                                                  # (nel, 4, 3)
    lo_idx = np.maximum(1, np.floor((tets.min(1) - g.AABB_min) / g.cell_size).astype(int) - 1)
    hi_idx = np.minimum(dims, np.ceil((tets.max(1) - g.AABB_min) / g.cell_size).astype(int) + 1)
    mism = checked = 0
    for k in range(int(dims[2])):
        for j in range(int(dims[1])):
            for i in range(int(dims[0])):
                x = g.AABB_min + g.cell_size * np.array([i, j, k])
                gi = np.clip(np.floor((x - g.AABB_min) / g.cell_size).astype(int) + 1, 1, dims)
                cand = np.where(((lo_idx <= gi) & (gi <= hi_idx)).all(1))[0]
                sign = -1.0
                for e in cand:
                    T = tets[e]
                    if (x < T.min(0) - 1e-10).any() or (x > T.max(0) + 1e-10).any():
                        continue
                    lam = np.linalg.solve(np.vstack([T.T, np.ones(4)]), np.append(x, 1.0))
                    if not ((lam >= -1e-10).all() and (lam <= 1 + 1e-10).all()):
                        continue
                    l234 = np.linalg.solve((T[1:] - T[0]).T, x - T[0])
                    l1 = 1.0 - l234.sum()
                    full = np.array([l1, *l234])
                    if not ((full >= 0).all() and full.sum() <= 1.0):
                        continue
                    N = np.array([l1, l234[0], l234[1], 1.0 - l1 - l234[0] - l234[1]])
                    if N @ rn[IEN[e] - 1] >= 0.5:
                        sign = 1.0
                        break
                checked += 1
                mism += int(sign != so[k, j, i])
    # LAPACK's pivoted solves and the oracle's adjugate formulas differ in the last bit: a point exactly on a shared face can be
    # accepted by one and rejected by the other (the reference's own sum(lambda) <= 1.0 test is that sharp) -- allow a handful
    assert checked == int(np.prod(dims)) and (so > 0).sum() > 50
    assert mism <= 0.01 * checked, (mism, checked)


def test_volume_from_sdf_matches_literal_restatement():
    """calculate_volume_from_sdf (CalcVolumeFromSDF.jl:26-125) restated literally in numpy (Float32 lerps in the reference's order, 9^3
    Gauss points in cut cells, full cells counted whole) on a random smooth field; and LS_Threshold's bisection (RBFs4Smoothing.jl:265-300)
    driven by it reaches the same offset as a bisection driven by the oracle's volume."""
    f32 = np.float32
    rng = np.random.default_rng(3)
    from scipy import ndimage
    sdf = ndimage.gaussian_filter(rng.standard_normal((11, 13, 12)), 1.5).astype(f32) * f32(4)       # [k, j, i]
    edge = f32(0.37)
    gp, gw = np.polynomial.legendre.leggauss(9)
    gp, gw = gp.astype(f32), gw.astype(f32)

    def volume(field, iso=f32(0)):
        c = {(a, b, d): field[d:field.shape[0] - 1 + d, b:field.shape[1] - 1 + b, a:field.shape[2] - 1 + a] for a in (0, 1) for b in (0, 1) for d in (0, 1)}   # c[(i,j,k) offsets]
        allv = np.stack(list(c.values()))
        mn, mx = allv.min(0), allv.max(0)
        ev = edge * edge * edge
        jac = ev / f32(8)
        total = np.float64((mn >= iso).sum()) * np.float64(ev)
        cut = (mx >= iso) & (mn < iso)
        part = np.zeros(cut.sum(), f32)
        cc = {k: v[cut] for k, v in c.items()}
        for kq in range(9):
            zeta = (gp[kq] + f32(1)) / f32(2)
            for jq in range(9):
                eta = (gp[jq] + f32(1)) / f32(2)
                for iq in range(9):
                    xi = (gp[iq] + f32(1)) / f32(2)
                    c00 = cc[(0, 0, 0)] * (f32(1) - xi) + cc[(1, 0, 0)] * xi
                    c01 = cc[(0, 0, 1)] * (f32(1) - xi) + cc[(1, 0, 1)] * xi
                    c10 = cc[(0, 1, 0)] * (f32(1) - xi) + cc[(1, 1, 0)] * xi
                    c11 = cc[(0, 1, 1)] * (f32(1) - xi) + cc[(1, 1, 1)] * xi
                    c0 = c00 * (f32(1) - eta) + c10 * eta
                    c1 = c01 * (f32(1) - eta) + c11 * eta
                    ps = c0 * (f32(1) - zeta) + c1 * zeta
                    part = np.where(ps >= iso, part + (gw[iq] * gw[jq] * gw[kq]) * jac, part).astype(f32)
        return float(total + np.float64(part.sum(dtype=np.float64)))

    v_lit = volume(sdf)
    v_or = oracle.volume_from_sdf(sdf, edge)
    assert abs(v_lit - v_or) <= 2e-6 * v_lit                       # Float32 summation order is the only freedom
    # LS_Threshold: th in [min, max], 40 steps, stop when |V_target - V| <= 1e-4 (absolute), V > target -> th_low = th
    target = 0.35 * float(np.prod(np.array(sdf.shape) - 1)) * float(edge) ** 3

    def bisect(vol):
        lo, hi, n, eps, th = f32(sdf.min()), f32(sdf.max()), 0, 1.0, f32(0)
        while n < 40 and eps > 1e-4:
            th = (lo + hi) / f32(2)
            cur = vol((sdf - th).astype(f32))
            eps = abs(target - cur)
            if cur > target:
                lo = th
            else:
                hi = th
            n += 1
        return float(th)

    th_lit = bisect(volume)
    th_or = bisect(lambda f: oracle.volume_from_sdf(f, edge))
    assert abs(th_lit - th_or) <= 1e-5 * float(np.abs(sdf).max())


def _dense_in_nodes_literal(X, IEN, rho):
    """DenseInNodes with FilterForNodalDensity, NodalDensityLeastSquares and LamReduction (NodalDensities.jl:89-218) in numpy, with
    LAPACK's symmetric eigen-solver standing where Julia calls eigen(A'A)."""
    nnp, nel = X.shape[0], IEN.shape[0]
    centre = X[IEN - 1].mean(axis=1)
    ine = [[] for _ in range(nnp)]
    for e in range(nel):
        for nd in IEN[e]:
            ine[nd - 1].append(e)
    out = np.zeros(nnp)
    for i in range(nnp):
        els = ine[i]
        if len(els) == 1:
            out[i] = rho[els[0]]
        elif len(els) < 4:
            L = np.array([np.linalg.norm(X[i] - centre[e]) for e in els])
            Lmax = L.max() * 1.2
            wts = 1 - L / Lmax
            out[i] = float((rho[els] * wts).sum() / wts.sum())
        else:
            A = np.column_stack([np.ones(len(els)), centre[els]])
            b = rho[els]
            lam, phi = np.linalg.eigh(A.T @ A)
            e1, e2, e3 = abs(lam.max() / lam.min()), abs(lam.max() / lam[1]), abs(lam.max() / lam[2])
            if 1e7 > e1 and 3e3 > e2:
                keep = lam
            elif 1e7 < e1 and 3e3 > e2:
                keep = lam[1:]
            elif 1e7 < e1 and 3e3 < e2:
                keep = lam[2:] if 3e3 > e3 else lam[3:]
            else:
                keep = np.array([])
            if keep.size == 0:
                out[i] = b.mean()
            else:
                poz = lam.size - keep.size
                b1 = phi.T @ (A.T @ b)
                x2 = np.concatenate([np.zeros(poz), b1[poz:] / keep])
                out[i] = float(np.concatenate([[1.0], X[i]]) @ (phi @ x2))
    return out


@pytest.mark.parametrize("name", ["sphere", "cantilever_beam_vfrac_03"])
def test_nodal_densities_match_literal_restatement(name):
    X, IEN, rho = load_mesh(name)
    lit = _dense_in_nodes_literal(X, IEN, rho)
    orc = oracle.nodal_densities(X, IEN, rho)
    # LAPACK eigh vs the oracle's Jacobi sweeps: with A'A conditioned like 1e4..1e7 (coordinates up to 60 on the cantilever) the two
    # eigen-decompositions differ by ~cond * eps in the fitted value -- a few 1e-12, the same level at which Julia's LAPACK would differ
    assert np.abs(lit - orc).max() <= (1e-12 if name == "sphere" else 1e-11)
    if name == "sphere":
        assert abs(lit.mean() - 0.29490556408887564) <= 1e-12          # the reference's golden (HexSphereSdfTest.jl:27)


def test_mesh_volume_matches_literal_restatement():
    """calculate_mesh_volume (MeshVolume.jl:4-117): 3^3 Gauss points, |det J| weights; HEX8 on an unstructured mesh and the TET4
    cube-to-tet mapping with the reference's (1-xi)^2 (1-xi-eta)/8 factor."""
    gp, gw = np.polynomial.legendre.leggauss(3)
    SG = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], float)
    X, IEN, rho = load_mesh("chapadlo")
    sel = np.arange(0, IEN.shape[0], 37)                # every 37th element is plenty
    vol = np.zeros(sel.size)
    for k in range(3):
        for j in range(3):
            for i in range(3):
                xi = np.array([gp[i], gp[j], gp[k]])
                t = 1 + SG * xi
                dN = 0.125 * np.stack([SG[:, 0] * t[:, 1] * t[:, 2], SG[:, 1] * t[:, 0] * t[:, 2], SG[:, 2] * t[:, 0] * t[:, 1]], axis=1)
                J = np.einsum("eai,aj->eij", X[IEN[sel] - 1], dN)
                vol += gw[i] * gw[j] * gw[k] * np.abs(np.linalg.det(J))
    vd, vf = oracle.mesh_volume(X, IEN[sel], rho[sel])
    assert abs(vd - vol.sum()) <= 1e-10 * vol.sum() and abs(vf - (vol * rho[sel]).sum() / vol.sum()) <= 1e-12
    # TET4
    Xt, T, rn = schlafli_tet4(3, "radial")
    dNt = np.array([[1, 0, 0], [0, 1, 0], [0, 0, 1], [-1, -1, -1]], float)
    vt = 0.0
    for e in range(T.shape[0]):
        Jd = abs(np.linalg.det(Xt[T[e] - 1].T @ dNt))
        for k in range(3):
            for j in range(3):
                for i in range(3):
                    xi = (gp[i] + 1) / 2; eta = (gp[j] + 1) / 2 * (1 - xi)
                    vt += gw[i] * gw[j] * gw[k] * Jd * (1 - xi) ** 2 * (1 - xi - eta) / 8.0
    vdt, _ = oracle.mesh_volume(Xt, T, np.ones(T.shape[0]))
    assert abs(vdt - vt) <= 1e-10 * vt and abs(vdt - 0.75 * 27) < 1e-9
