"""Shared test fixtures: reference meshes (tests/golden/*.npz), the block primitive, synthetic SIMP fields."""
import math
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class Grid:
    """Mirror of MeshGrid.Grid (reference src/MeshGrid/Grid.jl:2-35), used by oracle-side tests.
    The product package has its own copy (host mirror of the Julia struct)."""

    def __init__(self, AABB_min, AABB_max, N_max, margineCells=3):
        amin = np.asarray(AABB_min, dtype=np.float64).copy()
        amax = np.asarray(AABB_max, dtype=np.float64).copy()
        cell = float(np.max(amax - amin) / N_max)
        amin = amin - margineCells * cell
        amax = amax + margineCells * cell
        N = np.ceil((amax - amin) / cell).astype(np.int64)
        amax = amin + N * cell
        self.AABB_min, self.AABB_max, self.N, self.cell_size = amin, amax, N, cell
        self.ngp = int(np.prod(N + 1))


def load_mesh(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return d["X"], d["IEN"], d["rho"]


def block_geometry(N):
    """TestGeometryBlock (reference src/PrimitiveGeometries/PrimitiveGeometries.jl:157-214)."""
    N = np.asarray(N, dtype=np.int64)
    side = 2.0
    delta = side / N.max()
    L = delta * N
    nn = int(np.prod(N + 1))
    X = np.zeros((nn, 3))
    nid = {}
    for i in range(N[0] + 1):
        for j in range(N[1] + 1):
            for k in range(N[2] + 1):
                n = i * (N[2] + 1) * (N[1] + 1) + j * (N[2] + 1) + k
                X[n] = [-L[0] / 2 + i * delta, -L[1] / 2 + j * delta, -L[2] / 2 + k * delta]
                nid[(i, j, k)] = n + 1
    IEN = np.zeros((int(np.prod(N)), 8), dtype=np.int64)
    rho = np.zeros(int(np.prod(N)))
    for i in range(N[0]):
        for j in range(N[1]):
            for k in range(N[2]):
                e = i * N[2] * N[1] + j * N[2] + k
                c = [(i, j, k), (i + 1, j, k), (i + 1, j + 1, k), (i, j + 1, k), (i, j, k + 1), (i + 1, j, k + 1), (i + 1, j + 1, k + 1), (i, j + 1, k + 1)]
                IEN[e] = [nid[t] for t in c]
                ctr = X[IEN[e] - 1].mean(axis=0)
                rho[e] = 1 - np.linalg.norm(ctr) / (math.sqrt(3) * side / 2)
    return X, IEN, rho


BLOCK_RHO_N = np.array([0.0, 0.0, 0.5, 0.5, 0.5, 0.5, 1.0, 1.0, 0.0, 0.0, 0.5, 0.5])  # test/HexBlockSdfTest.jl:55


def _splitmix64(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & np.uint64(0xFFFFFFFFFFFFFFFF)
    return z ^ (z >> np.uint64(31))


def simp_hex8(n, seed=20240517, period_frac=0.25, t=-0.6, beta=8.0, noise=0.02):
    """Synthetic HEX8 SIMP field of SURVEY.md section 8(d)-5 on n^3 unit elements (nodes on the integer lattice
    [0,n]^3, VTK node order), omega = 2*pi/(period_frac*n) (period 64 at n = 256).  Returns X (nnp,3), IEN (nel,8)
    1-based, rho (nel)."""
    m = n + 1
    k, j, i = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    X = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    ke, je, ie = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ke, je, ie = ke.ravel(), je.ravel(), ie.ravel()
    nid = lambda a, b, c: (c * m + b) * m + a + 1
    IEN = np.stack([nid(ie, je, ke), nid(ie + 1, je, ke), nid(ie + 1, je + 1, ke), nid(ie, je + 1, ke),
                    nid(ie, je, ke + 1), nid(ie + 1, je, ke + 1), nid(ie + 1, je + 1, ke + 1), nid(ie, je + 1, ke + 1)], axis=1).astype(np.int64)
    w = 2 * np.pi / (period_frac * n)
    cx, cy, cz = ie + 0.5, je + 0.5, ke + 0.5
    g = np.sin(w * cx) * np.cos(w * cy) + np.sin(w * cy) * np.cos(w * cz) + np.sin(w * cz) * np.cos(w * cx)
    with np.errstate(over="ignore"):
        h = _splitmix64(np.arange(n ** 3, dtype=np.uint64) + np.uint64(seed))
    u = (h >> np.uint64(11)).astype(np.float64) / float(1 << 53) * 2.0 - 1.0
    rho = np.clip(1.0 / (1.0 + np.exp(-beta * (t - g))) + noise * u, 0.0, 1.0)
    return X, IEN, rho


def lattice_hex8(n):
    """n^3 unit HEX8 elements on the integer lattice [0,n]^3 (VTK node order); X (nnp,3), IEN (nel,8) 1-based."""
    m = n + 1
    k, j, i = np.meshgrid(np.arange(m), np.arange(m), np.arange(m), indexing="ij")
    X = np.stack([i.ravel(), j.ravel(), k.ravel()], axis=1).astype(np.float64)
    ke, je, ie = np.meshgrid(np.arange(n), np.arange(n), np.arange(n), indexing="ij")
    ke, je, ie = ke.ravel(), je.ravel(), ie.ravel()
    nid = lambda a, b, c: (c * m + b) * m + a + 1
    IEN = np.stack([nid(ie, je, ke), nid(ie + 1, je, ke), nid(ie + 1, je + 1, ke), nid(ie, je + 1, ke),
                    nid(ie, je, ke + 1), nid(ie + 1, je, ke + 1), nid(ie + 1, je + 1, ke + 1), nid(ie, je + 1, ke + 1)], axis=1).astype(np.int64)
    return X, IEN


def graded_lattice_hex8(n, seed=3, holes=0.12):
    """A tensor-product lattice mesh that is NOT the plain cube: graded (non-uniform) spacing per axis, ~12 % of the cells missing,
    element and node numbering shuffled.  Returns X, IEN (1-based), rho (element densities of the SIMP field at the cell centres)."""
    rng = np.random.default_rng(seed)
    X0, IEN0, rho0 = simp_hex8(n)
    ax = [np.concatenate([[0.0], np.cumsum(rng.uniform(0.6, 1.5, n))]) for _ in range(3)]
    X = np.stack([ax[0][X0[:, 0].astype(int)], ax[1][X0[:, 1].astype(int)], ax[2][X0[:, 2].astype(int)]], axis=1)
    keep = rng.random(IEN0.shape[0]) >= holes
    order = rng.permutation(np.nonzero(keep)[0])
    IEN, rho = IEN0[order], rho0[order]
    used = np.unique(IEN)
    perm = rng.permutation(used.size)
    new_id = np.zeros(X.shape[0] + 1, dtype=np.int64)
    new_id[used] = perm + 1
    Xn = np.zeros((used.size, 3)); Xn[perm] = X[used - 1]
    return Xn, np.ascontiguousarray(new_id[IEN]), rho


SCHLAFLI = np.array([[1, 2, 3, 7], [1, 6, 2, 7], [1, 3, 4, 7], [1, 4, 8, 7], [1, 5, 6, 7], [1, 8, 5, 7]]) - 1   # SimpleCubeWithSchlafli.jl:22-29


def schlafli_tet4(n, field="radial"):
    """Schlaefli 6-tet split of the n^3 lattice cube (reference test/PrimitiveGeometriesTest/SimpleCubeWithSchlafli.jl:22-29)
    with a nodal density: 'radial' = linear radial fall-off from the cube centre (the reference generators' field),
    'simp' = the synthetic SIMP field of SURVEY.md 8(d) sampled at the nodes.  Returns X, IEN (6 n^3, 4) 1-based, rho_n."""
    X, H = lattice_hex8(n)
    T = np.ascontiguousarray(np.concatenate([H[:, s] for s in SCHLAFLI], axis=0))
    if field == "radial":
        r = np.linalg.norm(X - n / 2.0, axis=1)
        rho_n = np.clip(1.0 - r / (0.75 * n), 0.0, 1.0)
    else:
        w = 2 * np.pi / (0.5 * n)
        g = np.sin(w * X[:, 0]) * np.cos(w * X[:, 1]) + np.sin(w * X[:, 1]) * np.cos(w * X[:, 2]) + np.sin(w * X[:, 2]) * np.cos(w * X[:, 0])
        rho_n = 1.0 / (1.0 + np.exp(-4.0 * (-0.6 - g)))
    return X, T, rho_n
