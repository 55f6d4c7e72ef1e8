"""Scale probe: synthetic n^3 HEX8 SIMP field -> (2n+...)^3 fine SDF through the resident pipeline. Usage: scale_probe.py n [reps]"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import rho2sdf_b200 as r2s
from fixtures import simp_hex8

n = int(sys.argv[1]); reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
t = time.time(); X, IEN, rho = simp_hex8(n); print("mesh gen %.1fs nel=%d" % (time.time() - t, IEN.shape[0]), flush=True)
t = time.time(); mesh = r2s.Mesh(X, IEN, rho); print("Mesh (upload, INE, faces, volume) %.2fs V_frac=%.4f" % (time.time() - t, mesh.V_frac), flush=True)
grid = r2s.Grid(X.min(0), X.max(0), 2 * n, 3); print("grid N=%s ngp=%d cell=%g" % (grid.N, grid.ngp, grid.cell_size), flush=True)
t = time.time(); rho_n = r2s.DenseInNodes(mesh, rho); print("DenseInNodes %.2fs" % (time.time() - t), flush=True)
mesh._use_grid(grid)
c = mesh.ctx
p = r2s.Params(); c.lib.r2s_default_params(C.byref(p))
p.rho_t = 0.5; p.smooth = 2; p.rbf_interp = 1; p.target_volume = mesh.V_frac * mesh.V_domain; p.final_volume = 1
c.check(c.lib.r2s_upload_nodal_densities(c.h, rho_n.ctypes.data_as(C.c_void_p)))
for r in range(reps):
    rep = r2s.Report(); t = time.time()
    c.check(c.lib.r2s_pipeline_resident(c.h, C.byref(p), C.byref(rep))); wall = time.time() - t
    d = rep.asdict(); nf = int(np.prod(grid.N * 2 + 1))
    print("rep %d wall %.3fs total %.1f ms -> %.3e fine voxels/s | bin %.1f proj %.1f asm %.1f sign %.1f cc %.1f prep %.1f cg %.1f (%d it) lsf %.1f thr %.1f (%d bis) fine %.1f vol %.1f | pairs=%d iters=%d notconv=%d flipped=%d launches=%d th=%.5f vol=%.1f"
          % (r, wall, d["ms_total"], nf / (d["ms_total"] * 1e-3), d["ms_bin"], d["ms_project"], d["ms_assemble"], d["ms_sign"], d["ms_cc"], d["ms_rbf_prep"], d["ms_cg"], d["cg_iters"],
             d["ms_lsf"], d["ms_threshold"], d["bisections"], d["ms_fine"], d["ms_volume"], d["n_pairs"], d["n_newton_iters"], d["n_not_converged"], d["n_flipped"], d["launches"], d["th"], d["volume"]), flush=True)
