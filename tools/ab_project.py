#!/usr/bin/env python
"""Side-by-side timing of the kernel variants in ONE process (the R2S_PROJ* / R2S_STENCIL knobs are read on every call).

    python tools/ab_project.py [--n 256] [--steps 2] [--variants "name:K=V,K=V;name2:..."]

Builds the bench workload once, then for every variant: sets the knobs, runs 1 warm-up + `steps` timed passes of
r2s_pipeline_resident and prints the stage times, the iteration counters and max |sdf - sdf(first variant)| / h (the first
variant is the reference the others are compared with).  Output: one JSON line per variant on stdout."""
import argparse, ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

DEFAULT = ("general:R2S_PROJ_BOX=0;box5:;box5_fast:R2S_PROJ_FAST=1;box5_p1:R2S_PROJ_P1=1;box5_fast_p1:R2S_PROJ_FAST=1,R2S_PROJ_P1=1;"
           "box5_uni:R2S_PROJ_UNI=1;box5_uni_p1:R2S_PROJ_UNI=1,R2S_PROJ_P1=1;box6_uni_p1:R2S_PROJ_UNI=1,R2S_PROJ_P1=1,R2S_PROJ_BOX_MINB=6;"
           "box8_uni_p1:R2S_PROJ_UNI=1,R2S_PROJ_P1=1,R2S_PROJ_BOX_MINB=8;refill3_box:R2S_PROJ=1;refill3_box_uni:R2S_PROJ=1,R2S_PROJ_UNI=1;"
           "refill4_box_uni:R2S_PROJ=1,R2S_PROJ_UNI=1,R2S_PROJ_MINB=4;general_fast_p1:R2S_PROJ_BOX=0,R2S_PROJ_FAST=1,R2S_PROJ_P1=1;general_uni_p1:R2S_PROJ_BOX=0,R2S_PROJ_UNI=1,R2S_PROJ_P1=1;"
           "box5_scaled:R2S_PROJ_SCALED=1;box5_scaled_p1:R2S_PROJ_SCALED=1,R2S_PROJ_P1=1;box6_scaled_p1:R2S_PROJ_SCALED=1,R2S_PROJ_P1=1,R2S_PROJ_BOX_MINB=6;"
           "box8_scaled_p1:R2S_PROJ_SCALED=1,R2S_PROJ_P1=1,R2S_PROJ_BOX_MINB=8;box5_atom:R2S_PROJ_ATOM=1;box6_atom_scaled:R2S_PROJ_ATOM=1,R2S_PROJ_SCALED=1,R2S_PROJ_BOX_MINB=6;stencil_fact:R2S_STENCIL=3;sign_class:R2S_SIGN_CLASS=1")
KNOBS = ("R2S_PROJ", "R2S_PROJ_BOX", "R2S_PROJ_BOX_MINB", "R2S_PROJ_MINB", "R2S_PROJ_SMEMA", "R2S_PROJ_FAST", "R2S_PROJ_UNI", "R2S_PROJ_SCALED", "R2S_PROJ_P1", "R2S_PROJ_ATOM", "R2S_STENCIL", "R2S_SIGN_CLASS")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--variants", default=DEFAULT)
    args = ap.parse_args()
    import torch
    import rho2sdf_b200 as r2s
    from fixtures import simp_hex8
    n = args.n
    X, IEN, rho = simp_hex8(n)
    mesh = r2s.Mesh(X, IEN, rho, element_type=r2s.HEX8)
    grid = r2s.Grid(X.min(0), X.max(0), 2 * n, 3)
    rho_n = r2s.DenseInNodes(mesh, rho)
    mesh._use_grid(grid)
    c = mesh.ctx
    p = r2s.Params(); c.lib.r2s_default_params(C.byref(p))
    p.rho_t, p.smooth, p.rbf_interp, p.remove_artifacts = 0.5, 2, 1, 1
    p.target_volume, p.final_volume = mesh.V_frac * mesh.V_domain, 1
    c.check(c.lib.r2s_upload_nodal_densities(c.h, rho_n.ctypes.data_as(C.c_void_p)))
    ref = None
    buf = torch.empty(int(grid.ngp), dtype=torch.float64).pin_memory()
    for spec in args.variants.split(";"):
        name, _, kv = spec.partition(":")
        for k in KNOBS:
            os.environ.pop(k, None)
        for item in filter(None, kv.split(",")):
            k, _, v = item.partition("=")
            os.environ[k] = v
        reps = []
        for it in range(1 + args.steps):
            rep = r2s.Report()
            c.check(c.lib.r2s_pipeline_resident(c.h, C.byref(p), C.byref(rep)))
            if it:
                reps.append(rep)
        c.check(c.lib.r2s_download_sdf(c.h, C.c_void_p(buf.data_ptr())))
        sdf = buf.numpy()
        if ref is None:
            ref = sdf.copy(); diff = 0.0; sign_diff = 0
        else:
            diff = float(np.max(np.abs(sdf - ref))) / grid.cell_size; sign_diff = int(np.count_nonzero(np.signbit(sdf) != np.signbit(ref)))
        out = {"variant": name, "knobs": kv, "n": n}
        for k in ("ms_bin", "ms_project", "ms_assemble", "ms_sign", "ms_cg", "ms_lsf", "ms_total"):
            out[k] = round(float(np.mean([getattr(r, k) for r in reps])), 3)
        out.update(pairs=int(reps[-1].n_pairs), newton_iters=int(reps[-1].n_newton_iters), not_converged=int(reps[-1].n_not_converged), cg_iters=int(reps[-1].cg_iters), th=float(reps[-1].th),
                   max_diff_over_h=diff, sign_diff=sign_diff)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
