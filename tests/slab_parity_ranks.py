"""Multi-rank slab parity (run under torchrun on N GPUs): the z-slab pipeline on N ranks against the single-rank pipeline run on
each rank's own GPU.  Distances, signs and the artifact removal must be bit-identical; the Float32 smoothing differs only by
the summation order of the CG dot products (tolerance 1e-5 * |field|).  Prints 'SLAB PARITY OK' on rank 0."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import rho2sdf_b200 as r2s
    from fixtures import simp_hex8
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 48
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    X, IEN, rho = simp_hex8(n)
    mesh = r2s.Mesh(X, IEN, rho, device=local)
    grid = r2s.Grid(X.min(0), X.max(0), 2 * n, 3)
    rn = r2s.DenseInNodes(mesh, rho)
    mesh._use_grid(grid)
    c = mesh.ctx
    # a noisy sign pattern would need a noisy density; instead lower the artifact ratio so that small components exist and get flipped
    p = r2s.Params(); c.lib.r2s_default_params(C.byref(p))
    p.rho_t, p.smooth, p.rbf_interp, p.remove_artifacts, p.artifact_min_ratio = 0.5, 2, 1, 1, 0.3
    p.target_volume, p.final_volume = mesh.V_frac * mesh.V_domain, 1
    nx, ny, nz = (int(v) + 1 for v in grid.N)
    fd = [int(v) * 2 + 1 for v in grid.N]

    def run():
        rep = r2s.Report()
        c.check(c.lib.r2s_upload_nodal_densities(c.h, rn.ctypes.data_as(C.c_void_p)))
        c.check(c.lib.r2s_pipeline_resident(c.h, C.byref(p), C.byref(rep)))
        sdf = np.empty(grid.ngp); c.check(c.lib.r2s_download_sdf(c.h, sdf.ctypes.data_as(C.c_void_p)))
        fine = np.empty(fd[0] * fd[1] * fd[2], dtype=np.float32); c.check(c.lib.r2s_download_fine_sdf(c.h, fine.ctypes.data_as(C.c_void_p)))
        return sdf.reshape(nz, ny, nx), fine.reshape(fd[2], fd[1], fd[0]), rep

    cases = [(0.5, 0.3), (0.93, 0.5), (0.08, 0.5)]                # (rho_t, artifact ratio): the extreme thresholds fragment the interior mask
    single = []
    for rt, ratio in cases:                                      # single rank, whole grid
        p.rho_t, p.artifact_min_ratio = rt, ratio
        single.append(run())
    k0, k1 = r2s.slab_partition(nz, world)[rank]
    r2s.init_slab_comm(c, rank, world, k0, k1)
    kf0, kf1 = 2 * k0, (2 * k1 if k1 < nz else fd[2])
    ok = True
    def check(name, cond):
        nonlocal ok
        if not cond:
            ok = False
            print("[rank %d] FAIL %s" % (rank, name), flush=True)
    summary = []
    for (rt, ratio), (sdf1, fine1, rep1) in zip(cases, single):
        p.rho_t, p.artifact_min_ratio = rt, ratio
        sdfN, fineN, repN = run()                               # this rank's slab of the N-rank run
        tag = "[rho_t=%g] " % rt
        check(tag + "sdf bit-identical on owned planes", np.array_equal(sdf1[k0:k1], sdfN[k0:k1]))
        check(tag + "flipped equal (%d vs %d)" % (rep1.n_flipped, repN.n_flipped), rep1.n_flipped == repN.n_flipped)
        check(tag + "cg iterations equal", rep1.cg_iters == repN.cg_iters)
        check(tag + "bisections equal", rep1.bisections == repN.bisections)
        scale = max(1.0, float(np.abs(fine1).max()))
        err = float(np.abs(fine1[kf0:kf1] - fineN[kf0:kf1]).max())
        check(tag + "fine field within 1e-5 (err %.3e, scale %.3g)" % (err, scale), err <= 1e-5 * scale)
        check(tag + "threshold offset", abs(rep1.th - repN.th) <= 1e-5 * scale)
        check(tag + "final volume", abs(rep1.volume - repN.volume) <= 1e-5 * abs(rep1.volume))
        check(tag + "collectives issued", repN.collectives > 0 and rep1.collectives == 0)
        summary.append("rho_t=%g flipped=%d cg=%d bis=%d th=%.7f/%.7f fine err=%.3e coll=%d ms1=%.2f msN=%.2f" %
                       (rt, repN.n_flipped, repN.cg_iters, repN.bisections, rep1.th, repN.th, err, repN.collectives, rep1.ms_total, repN.ms_total))
    p.rho_t, p.artifact_min_ratio = cases[0]
    sdfN, fineN, repN = run()
    # host-buffer slab entry point returns exactly the owned planes
    h_sdf = np.empty((k1 - k0) * ny * nx); h_fine = np.empty((kf1 - kf0) * fd[1] * fd[0], dtype=np.float32); rep = r2s.Report()
    c.check(c.lib.r2s_pipeline_slab(c.h, C.byref(p), rn.ctypes.data_as(C.c_void_p), h_sdf.ctypes.data_as(C.c_void_p), h_fine.ctypes.data_as(C.c_void_p), C.byref(rep)))
    check("pipeline_slab sdf", np.array_equal(h_sdf.reshape(k1 - k0, ny, nx), sdfN[k0:k1]))
    check("pipeline_slab fine", np.array_equal(h_fine.reshape(kf1 - kf0, fd[1], fd[0]), fineN[kf0:kf1]))
    # a LARGER grid on the live communicator: every field is re-allocated, the peers' mappings of the CG vector must follow
    grid2 = r2s.Grid(X.min(0), X.max(0), 3 * n, 3)
    nx2, ny2, nz2 = (int(v) + 1 for v in grid2.N)
    fd2 = [int(v) * 2 + 1 for v in grid2.N]
    ref = r2s.Mesh(X, IEN, rho, device=local)                      # single-rank reference on its own context
    ref._use_grid(grid2)
    rc_ = ref.ctx
    p.rho_t, p.artifact_min_ratio = 0.5, 0.3
    rep1 = r2s.Report()
    rc_.check(rc_.lib.r2s_upload_nodal_densities(rc_.h, rn.ctypes.data_as(C.c_void_p)))
    rc_.check(rc_.lib.r2s_pipeline_resident(rc_.h, C.byref(p), C.byref(rep1)))
    sdf1 = np.empty(grid2.ngp); rc_.check(rc_.lib.r2s_download_sdf(rc_.h, sdf1.ctypes.data_as(C.c_void_p)))
    fine1 = np.empty(fd2[0] * fd2[1] * fd2[2], dtype=np.float32); rc_.check(rc_.lib.r2s_download_fine_sdf(rc_.h, fine1.ctypes.data_as(C.c_void_p)))
    ref.ctx.close()
    mesh._use_grid(grid2)                                           # r2s_set_grid resets the slab ...
    k0, k1 = r2s.slab_partition(nz2, world)[rank]
    c.check(c.lib.r2s_set_slab(c.h, k0, k1))                        # ... so it is set again (collective)
    repN = r2s.Report()
    c.check(c.lib.r2s_upload_nodal_densities(c.h, rn.ctypes.data_as(C.c_void_p)))
    c.check(c.lib.r2s_pipeline_resident(c.h, C.byref(p), C.byref(repN)))
    sdfN = np.empty(grid2.ngp); c.check(c.lib.r2s_download_sdf(c.h, sdfN.ctypes.data_as(C.c_void_p)))
    fineN = np.empty(fd2[0] * fd2[1] * fd2[2], dtype=np.float32); c.check(c.lib.r2s_download_fine_sdf(c.h, fineN.ctypes.data_as(C.c_void_p)))
    pl2, fpl2 = nx2 * ny2, fd2[0] * fd2[1]
    kf0, kf1 = 2 * k0, (2 * k1 if k1 < nz2 else fd2[2])
    check("[grown grid] sdf bit-identical", np.array_equal(sdf1[k0 * pl2:k1 * pl2], sdfN[k0 * pl2:k1 * pl2]))
    check("[grown grid] cg iterations", rep1.cg_iters == repN.cg_iters)
    scale = max(1.0, float(np.abs(fine1).max()))
    check("[grown grid] fine field", float(np.abs(fine1[kf0 * fpl2:kf1 * fpl2] - fineN[kf0 * fpl2:kf1 * fpl2]).max()) <= 1e-5 * scale)
    t = torch.tensor([1 if ok else 0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        for ln in summary:
            print(ln, flush=True)
        print("SLAB PARITY OK" if int(t.item()) == 1 else "SLAB PARITY FAILED", flush=True)
    dist.barrier()
    mesh.ctx.close()
    dist.destroy_process_group()
    return 0 if int(t.item()) == 1 else 1


if __name__ == "__main__":
    sys.exit(main())
