"""ctypes binding of the CPU oracle (oracle/r2s_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and the cpu_baseline /
--impl reference legs of bench.py.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_i64 = C.c_int64
_dp = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_ip = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_fp = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "libr2s_oracle.so")
    src = os.path.join(_HERE, "r2s_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libr2s_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.r2so_mesh_volume.argtypes = [c_i64, _dp, c_i64, C.c_int, _ip, _dp, C.POINTER(C.c_double), C.POINTER(C.c_double)]
        L.r2so_nodal_densities.argtypes = [c_i64, _dp, c_i64, C.c_int, _ip, _dp, _dp]
        L.r2so_isocontour_volume.argtypes = [c_i64, _dp, c_i64, _ip, _dp, C.c_double]
        L.r2so_isocontour_volume.restype = C.c_double
        L.r2so_find_threshold.argtypes = [c_i64, _dp, c_i64, _ip, _dp, C.c_double, C.c_double, C.c_int, C.POINTER(C.c_double)]
        L.r2so_eval_distances.argtypes = [c_i64, _dp, c_i64, C.c_int, _ip, _dp, _dp, _ip, C.c_double, _dp, C.c_double, C.c_double,
                                          C.c_int, _dp, C.c_void_p, C.c_void_p]
        L.r2so_sign_detection.argtypes = [c_i64, _dp, c_i64, C.c_int, _ip, _dp, _dp, _ip, C.c_double, _dp, C.c_double, C.c_int, _dp]
        L.r2so_remove_artifacts.argtypes = [_dp, _ip, C.c_double, C.c_double, C.POINTER(c_i64)]
        L.r2so_volume_from_sdf.argtypes = [_fp, c_i64, c_i64, c_i64, C.c_float, C.c_float, C.c_int, C.c_int]
        L.r2so_volume_from_sdf.restype = C.c_double
        L.r2so_rbf_smoothing.argtypes = [_dp, _dp, _dp, _ip, C.c_double, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, C.c_int,
                                         _fp, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_int), C.c_void_p, C.c_void_p]
        L.r2so_project_iso_hex8.argtypes = [_dp, C.c_double, _dp, _dp, _dp, C.POINTER(C.c_int)]
        L.r2so_inverse_map_hex8.argtypes = [_dp, _dp, _dp]
        L.r2so_project_iso_tet4.argtypes = [_dp, C.c_double, _dp, _dp, _dp]
        L.r2so_max_threads.restype = C.c_int
    return _LIB


def _mesh(X, IEN):
    X = np.ascontiguousarray(X, dtype=np.float64)      # (nnp,3) == Julia 3 x nnp
    IEN = np.ascontiguousarray(IEN, dtype=np.int64)    # (nel,nen) == Julia nen x nel, 1-based
    return X, IEN, X.shape[0], IEN.shape[0], IEN.shape[1]


def _grid(grid):
    return (np.ascontiguousarray(grid.AABB_min, dtype=np.float64), np.ascontiguousarray(grid.AABB_max, dtype=np.float64),
            np.ascontiguousarray(grid.N, dtype=np.int64), float(grid.cell_size))


def max_threads():
    return int(lib().r2so_max_threads())


def mesh_volume(X, IEN, rho):
    X, IEN, nnp, nel, nen = _mesh(X, IEN)
    vd, vf = C.c_double(), C.c_double()
    lib().r2so_mesh_volume(nnp, X, nel, nen, IEN, np.ascontiguousarray(rho, dtype=np.float64), C.byref(vd), C.byref(vf))
    return vd.value, vf.value


def nodal_densities(X, IEN, rho):
    X, IEN, nnp, nel, nen = _mesh(X, IEN)
    out = np.zeros(nnp)
    lib().r2so_nodal_densities(nnp, X, nel, nen, IEN, np.ascontiguousarray(rho, dtype=np.float64), out)
    return out


def isocontour_volume(X, IEN, rho_n, thr):
    X, IEN, nnp, nel, nen = _mesh(X, IEN)
    return lib().r2so_isocontour_volume(nnp, X, nel, IEN, np.ascontiguousarray(rho_n), float(thr))


def find_threshold(X, IEN, rho_n, target, tol=1e-4, maxit=60):
    X, IEN, nnp, nel, nen = _mesh(X, IEN)
    out = C.c_double()
    rc = lib().r2so_find_threshold(nnp, X, nel, IEN, np.ascontiguousarray(rho_n), float(target), tol, maxit, C.byref(out))
    if rc:
        raise RuntimeError("Requested volume is outside the possible range")
    return out.value


def eval_distances(X, IEN, grid, rho_n, rho_t, delta_factor=1.1, nthreads=1, want_xp=True):
    X, IEN, nnp, nel, nen = _mesh(X, IEN)
    amin, amax, N, cell = _grid(grid)
    ngp = int(np.prod(N + 1))
    dist = np.zeros(ngp)
    xp = np.zeros((ngp, 3)) if want_xp else None
    stats = np.zeros(3, dtype=np.int64)
    lib().r2so_eval_distances(nnp, X, nel, nen, IEN, amin, amax, N, cell, np.ascontiguousarray(rho_n, dtype=np.float64), float(rho_t),
                              float(delta_factor), int(nthreads), dist, xp.ctypes.data if want_xp else None, stats.ctypes.data)
    return dist, xp, {"pairs": int(stats[0]), "iters": int(stats[1]), "not_converged": int(stats[2])}


def sign_detection(X, IEN, grid, rho_n, rho_t, nthreads=1):
    X, IEN, nnp, nel, nen = _mesh(X, IEN)
    amin, amax, N, cell = _grid(grid)
    signs = np.zeros(int(np.prod(N + 1)))
    lib().r2so_sign_detection(nnp, X, nel, nen, IEN, amin, amax, N, cell, np.ascontiguousarray(rho_n, dtype=np.float64), float(rho_t), int(nthreads), signs)
    return signs


def remove_artifacts(sdf, grid, threshold=0.0, min_component_ratio=0.01):
    """In-place on a float64 copy; returns (sdf, flipped)."""
    sdf = np.ascontiguousarray(sdf, dtype=np.float64).copy()
    fl = c_i64()
    lib().r2so_remove_artifacts(sdf, np.ascontiguousarray(grid.N, dtype=np.int64), float(threshold), float(min_component_ratio), C.byref(fl))
    return sdf, int(fl.value)


def volume_from_sdf(sdf3, edge, iso=0.0, order=9, nthreads=1):
    """sdf3: float32 array indexed [k,j,i] (x fastest)."""
    a = np.ascontiguousarray(sdf3, dtype=np.float32)
    nz, ny, nx = a.shape
    return lib().r2so_volume_from_sdf(a, nx, ny, nz, np.float32(edge), np.float32(iso), order, nthreads)


def rbf_smoothing(sdf, grid, is_interp, smooth, target_volume, rbf_cut=1e-3, mode=0, nthreads=1, want_aux=False):
    amin, amax, N, cell = _grid(grid)
    dims = N * smooth + 1
    fine = np.zeros(int(np.prod(dims)), dtype=np.float32)
    th, vol, it = C.c_float(), C.c_float(), C.c_int()
    ngp = int(np.prod(N + 1))
    w = np.zeros(ngp, dtype=np.float32) if want_aux else None
    lsf = np.zeros(ngp, dtype=np.float32) if want_aux else None
    rc = lib().r2so_rbf_smoothing(np.ascontiguousarray(sdf, dtype=np.float64), amin, amax, N, cell, int(bool(is_interp)), int(smooth), float(rbf_cut),
                                  float(target_volume), int(mode), int(nthreads), fine, C.byref(th), C.byref(vol), C.byref(it),
                                  w.ctypes.data if want_aux else None, lsf.ctypes.data if want_aux else None)
    if rc:
        raise RuntimeError("rbf_smoothing failed rc=%d" % rc)
    info = {"th": th.value, "volume": vol.value, "cg_iters": it.value, "weights": w, "lsf": lsf}
    return fine.reshape(int(dims[2]), int(dims[1]), int(dims[0])), info


def project_iso_hex8(x, rho_t, Xe, re):
    """Xe: (8,3). Returns (ok, xi, iters)."""
    xi = np.zeros(3)
    it = C.c_int()
    ok = lib().r2so_project_iso_hex8(np.ascontiguousarray(x, dtype=np.float64), float(rho_t), np.ascontiguousarray(Xe, dtype=np.float64),
                                     np.ascontiguousarray(re, dtype=np.float64), xi, C.byref(it))
    return bool(ok), xi, it.value


def inverse_map_hex8(x, Xe):
    """Xe: (8,3). Returns (ok, xi)."""
    xi = np.zeros(3)
    ok = lib().r2so_inverse_map_hex8(np.ascontiguousarray(x, dtype=np.float64), np.ascontiguousarray(Xe, dtype=np.float64), xi)
    return bool(ok), xi


def project_iso_tet4(x, rho_t, Xe, re):
    """Xe: (4,3). Returns (ok, xp)."""
    xp = np.zeros(3)
    ok = lib().r2so_project_iso_tet4(np.ascontiguousarray(x, dtype=np.float64), float(rho_t), np.ascontiguousarray(Xe, dtype=np.float64),
                                     np.ascontiguousarray(re, dtype=np.float64), xp)
    return bool(ok), xp
