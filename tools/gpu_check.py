"""Diagnostic parity run (GPU vs oracle) with verbose statistics; not a test. Usage: python tools/gpu_check.py [cases...]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import oracle
import rho2sdf_b200 as r2s
from fixtures import load_mesh, block_geometry, BLOCK_RHO_N, simp_hex8


def cmp(name, a, b, h):
    d = np.abs(a - b)
    bad = d > 1e-9 * h
    print("   %-10s max|diff|/h = %.3e   n(>1e-9h) = %d / %d" % (name, d.max() / h, int(bad.sum()), a.size))
    return bad


def run(name, X, IEN, rho, rho_n, rho_t, grid, deltas=(1.1, 2.5), smooth=2, do_rbf=True):
    print("== %s: nel=%d nnp=%d grid N=%s ngp=%d h=%.4g" % (name, IEN.shape[0], X.shape[0], grid.N, grid.ngp, grid.cell_size))
    t = time.time(); mesh = r2s.Mesh(X, IEN, rho, element_type=r2s.HEX8); print("   Mesh upload+tables %.3fs  V_domain=%.10g V_frac=%.10g" % (time.time() - t, mesh.V_domain, mesh.V_frac))
    ovd, ovf = oracle.mesh_volume(X, IEN, rho); print("   oracle volume            V_domain=%.10g V_frac=%.10g" % (ovd, ovf))
    if rho_n is None:
        rho_n = r2s.DenseInNodes(mesh, rho); orn = oracle.nodal_densities(X, IEN, rho)
        print("   DenseInNodes: max|gpu-oracle| = %.3e  bit-identical=%s" % (np.abs(rho_n - orn).max(), np.array_equal(rho_n, orn)))
        rho_n = orn
    if rho_t is None:
        t = time.time(); rho_t = r2s.find_threshold_for_volume(mesh, rho_n); tg = time.time() - t
        t = time.time(); ort = oracle.find_threshold(X, IEN, rho_n, ovd * ovf); to = time.time() - t
        print("   threshold gpu=%.12g (%.2fs) oracle=%.12g (%.2fs)" % (rho_t, tg, ort, to)); rho_t = ort
    for df in deltas:
        t = time.time(); d, xp = r2s.evalDistances(mesh, grid, None, rho_n, rho_t, delta_factor=df); tg = time.time() - t
        rep = mesh.ctx.report()
        t = time.time(); od, oxp, st = oracle.eval_distances(X, IEN, grid, rho_n, rho_t, df); to = time.time() - t
        print("  delta=%.1f gpu %.3fs (bin %.2f ms, project %.2f ms, assemble %.2f ms; pairs=%d iters=%d notconv=%d active=%d)  oracle %.2fs (pairs=%d iters=%d notconv=%d)"
              % (df, tg, rep.ms_bin, rep.ms_project, rep.ms_assemble, rep.n_pairs, rep.n_newton_iters, rep.n_not_converged, rep.n_active, to, st["pairs"], st["iters"], st["not_converged"]))
        bad = cmp("dist", d, od, grid.cell_size)
        far = (od > 1e9)
        print("   far points gpu=%d oracle=%d" % (int((d > 1e9).sum()), int(far.sum())))
        if bad.any():
            idx = np.where(bad)[0][:8]
            for v in idx: print("     v=%d gpu=%.15g oracle=%.15g" % (v, d[v], od[v]))
        near = ~far
        cmp("xp", xp[near].ravel(), oxp[near].ravel(), grid.cell_size)
    s = r2s.Sign_Detection(mesh, grid, None, rho_n, rho_t); os_ = oracle.sign_detection(X, IEN, grid, rho_n, rho_t)
    print("   sign: mismatches=%d  npos gpu=%d oracle=%d" % (int((s != os_).sum()), int((s > 0).sum()), int((os_ > 0).sum())))
    sdf = od * os_
    g2 = sdf.copy(); nf = r2s.remove_sdf_artifacts(g2, grid, mesh=mesh); o2, onf = oracle.remove_artifacts(sdf, grid)
    print("   artifacts: flipped gpu=%d oracle=%d identical=%s" % (nf, onf, np.array_equal(g2, o2)))
    # a noisier mask to exercise the CC
    rng = np.random.default_rng(1); noisy = sdf.copy(); m = rng.random(sdf.size) < 0.03; noisy[m] = np.abs(noisy[m]) * np.where(rng.random(m.sum()) < 0.5, 1, -1)
    g3 = noisy.copy(); nf = r2s.remove_sdf_artifacts(g3, grid, mesh=mesh); o3, onf = oracle.remove_artifacts(noisy, grid)
    print("   artifacts(noisy): flipped gpu=%d oracle=%d identical=%s" % (nf, onf, np.array_equal(g3, o3)))
    if do_rbf:
        target = ovd * ovf
        for interp in (True, False):
            for sm in ((1, smooth) if smooth != 1 else (1,)):
                t = time.time(); fine, fg, info = r2s.RBFs_smoothing(mesh, o2, grid, interp, sm, name, return_info=True); tg = time.time() - t
                rep = mesh.ctx.report()
                t = time.time(); ofine, oinfo = oracle.rbf_smoothing(o2, grid, interp, sm, target, mode=1, nthreads=8); to = time.time() - t
                ofine0, oinfo0 = oracle.rbf_smoothing(o2, grid, interp, sm, target, mode=0, nthreads=8)
                dd = np.abs(fine - ofine).max() / grid.cell_size; dd0 = np.abs(fine - ofine0).max() / grid.cell_size; d10 = np.abs(ofine - ofine0).max() / grid.cell_size
                print("   rbf interp=%d smooth=%d: gpu %.3fs (cg %.2f ms/%d it, lsf %.2f, thr %.2f ms/%d bis, fine %.2f, vol %.2f) oracle %.2fs (%d it)"
                      % (interp, sm, tg, rep.ms_cg, info["cg_iters"], rep.ms_lsf, rep.ms_threshold, info["bisections"], rep.ms_fine, rep.ms_volume, to, oinfo["cg_iters"]))
                print("      max|gpu-oracle(ideal)|/h=%.3e  |gpu-oracle(faithful)|/h=%.3e  |ideal-faithful|/h=%.3e  th gpu=%.7g ideal=%.7g faithful=%.7g  vol gpu=%.6g ideal=%.6g target=%.6g"
                      % (dd, dd0, d10, info["th"], oinfo["th"], oinfo0["th"], info["volume"], oinfo["volume"], target))
    mesh.ctx.close()


if __name__ == "__main__":
    cases = sys.argv[1:] or ["block", "sphere", "cantilever", "simp16"]
    if "block" in cases:
        X, IEN, rho = block_geometry([2, 1, 1]); run("block", X, IEN, rho, BLOCK_RHO_N, 0.5, r2s.Grid(X.min(0), X.max(0), 20, 3))
    if "sphere" in cases:
        X, IEN, rho = load_mesh("sphere"); run("sphere", X, IEN, rho, None, 0.5, r2s.Grid(X.min(0), X.max(0), 10, 3))
    if "cantilever" in cases:
        X, IEN, rho = load_mesh("cantilever_beam_vfrac_03"); m = r2s.Mesh(X, IEN, rho); g = r2s.noninteractive_sdf_grid_setup(m); m.ctx.close()
        run("cantilever", X, IEN, rho, None, None, g, deltas=(1.1,))
    if "simp16" in cases:
        X, IEN, rho = simp_hex8(16); run("simp16", X, IEN, rho, None, 0.5, r2s.Grid(X.min(0), X.max(0), 32, 3), deltas=(1.1,))
    if "chapadlo" in cases:
        X, IEN, rho = load_mesh("chapadlo"); m = r2s.Mesh(X, IEN, rho); g = r2s.noninteractive_sdf_grid_setup(m); m.ctx.close()
        run("chapadlo", X, IEN, rho, None, None, g, deltas=(1.1,))
