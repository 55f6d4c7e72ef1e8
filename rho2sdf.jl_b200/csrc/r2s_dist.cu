// r2s_dist.cu -- evalDistances (SignedDistances/sdfOnDensityField.jl:139-486) on the GPU
//
//  1. binning   : classify elements (solid / crossing / void, :199-201,:312), give every active element the grid-point
//                 range of its AABB +- delta cell range (MeshGrid/Grid.jl:122-154) and bin it into the voxel tiles it
//                 overlaps: (tile, element) keys -> radix sort -> per-tile element lists in ascending element order.
//  2. project   : element-centric, one warp per 32 grid points of a crossing element: closest point on the in-element
//                 iso-surface (r2s_iso.cuh), distance written to a per-(element, point) pair buffer.  FP64-bound.
//  3. assemble  : voxel-centric, one thread per grid point: walks its tile's element list IN ELEMENT ORDER and replays
//                 the reference's running-min logic exactly (boundary-face triangles with their state-dependent
//                 early-outs, then the element's iso distance from the pair buffer).  Bit-exact decisions (r2s_exact.cuh).
//                 No atomics: every voxel is owned by one thread, the result is deterministic.
#include <stdlib.h>
#include <type_traits>
#include "r2s_common.cuh"
#include "r2s_tables.cuh"
#include "r2s_exact.cuh"
#include "r2s_iso.cuh"

// ------------------------------------------------------------------------------------------------ binning
// [zlo, zhi]: z-extent of the slab's planes grown by a safe margin; elements entirely outside cannot reach a plane of this rank
// and are left inactive without touching their nodes (single rank: the interval is infinite).  The class counters then
// count the elements near the slab only.
__global__ void k_classify(i64 nel, int nen, const int *__restrict__ IEN, const double *__restrict__ rn, const unsigned char *__restrict__ fb,
                           const double2 *__restrict__ ezr, double zlo, double zhi,
                           double rho_t, unsigned char *__restrict__ cls, int *__restrict__ flag, i64 *__restrict__ counts) {
  i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  int c = 0;
  if (e < nel) { const double2 z = ezr[e]; if (z.y < zlo || z.x > zhi) { cls[e] = 0; flag[e] = 0; e = nel; } }
  if (e < nel) {
    double mn = 1e300, mx = -1e300;
    for (int a = 0; a < nen; a++) { double r = rn[IEN[nen * e + a]]; mn = fmin(mn, r); mx = fmax(mx, r); }
    if (mn >= rho_t) c = 1; else if (mx > rho_t) c = 2;
    cls[e] = (unsigned char)c;
    flag[e] = (c == 2 || (c == 1 && fb[e] != 0)) ? 1 : 0;
  }
  // class counters (report only): one pair of atomics per CTA
  __shared__ int s1, s2;
  if (threadIdx.x == 0) { s1 = 0; s2 = 0; }
  __syncthreads();
  unsigned m1 = __ballot_sync(0xffffffffu, c == 1), m2 = __ballot_sync(0xffffffffu, c == 2);
  if ((threadIdx.x & 31) == 0) { if (m1) atomicAdd(&s1, __popc(m1)); if (m2) atomicAdd(&s2, __popc(m2)); }
  __syncthreads();
  if (threadIdx.x == 0) { if (s1) atomicAdd((u64 *)&counts[0], (u64)s1); if (s2) atomicAdd((u64 *)&counts[1], (u64)s2); }
}
// compacted active list: record with point ranges, tile count, pair count
__global__ void k_act_records(i64 nel, int nen, const int *__restrict__ IEN, const double *__restrict__ X, const unsigned char *__restrict__ cls,
                              const unsigned char *__restrict__ fb, const int *__restrict__ flag, const int *__restrict__ idx, GridDev g, double delta,
                              int kz0, int kz1, ActRec *__restrict__ rec, i64 *__restrict__ ntile, i64 *__restrict__ npair, i64 *__restrict__ nchunk) {
  i64 e = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (e >= nel || !flag[e]) return;
  double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
  for (int a = 0; a < nen; a++) { i64 n = IEN[nen * e + a]; for (int d = 0; d < 3; d++) { double c = X[3 * n + d]; lo[d] = fmin(lo[d], c); hi[d] = fmax(hi[d], c); } }
  ActRec r; r.el = (int)e; r.cls = cls[e]; r.fmask = fb[e]; r.tri_off = 0; r.pair_off = 0;
  bool ok = true;
  for (int d = 0; d < 3; d++) {
    int I0 = 0, I1 = -1;
    ok = ok && ex::cell_range_axis(lo[d], hi[d], delta, g.amin[d], g.amax[d], g.N[d], I0, I1);
    if (ok) { r.ps[d] = g.cstart[g.cs_off[d] + I0]; r.pe[d] = g.cstart[g.cs_off[d] + I1 + 1]; } else { r.ps[d] = 0; r.pe[d] = 0; }
  }
  // z-slab restriction (multi-GPU): only planes [kz0,kz1)
  if (r.ps[2] < kz0) r.ps[2] = kz0;
  if (r.pe[2] > kz1) r.pe[2] = kz1;
  i64 vol = 0, nt = 0;
  if (ok && r.pe[0] > r.ps[0] && r.pe[1] > r.ps[1] && r.pe[2] > r.ps[2]) {
    vol = (i64)(r.pe[0] - r.ps[0]) * (r.pe[1] - r.ps[1]) * (r.pe[2] - r.ps[2]);
    nt = (i64)((r.pe[0] - 1) / TILE_X - r.ps[0] / TILE_X + 1) * ((r.pe[1] - 1) / TILE_Y - r.ps[1] / TILE_Y + 1) * ((r.pe[2] - 1) / TILE_Z - r.ps[2] / TILE_Z + 1);
  } else { r.pe[0] = r.ps[0]; r.pe[1] = r.ps[1]; r.pe[2] = r.ps[2]; }
  int a = idx[e];
  rec[a] = r; ntile[a] = nt;
  i64 np = (r.cls == 2) ? vol : 0;
  npair[a] = np; nchunk[a] = (np + 31) / 32;
}
__global__ void k_emit_keys(i64 nact, ActRec *__restrict__ rec, const i64 *__restrict__ toff, const i64 *__restrict__ poff, GridDev g,
                            u64 *__restrict__ keys, int *__restrict__ tile_cnt, unsigned char *__restrict__ tile_faces) {
  i64 a = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (a >= nact) return;
  ActRec r = rec[a];
  rec[a].pair_off = poff[a];
  if (r.pe[0] <= r.ps[0]) return;
  i64 o = toff[a];
  for (int tz = r.ps[2] / TILE_Z; tz <= (r.pe[2] - 1) / TILE_Z; tz++)
    for (int ty = r.ps[1] / TILE_Y; ty <= (r.pe[1] - 1) / TILE_Y; ty++)
      for (int tx = r.ps[0] / TILE_X; tx <= (r.pe[0] - 1) / TILE_X; tx++) {
        u64 t = ((u64)tz * g.nt[1] + ty) * g.nt[0] + tx;
        keys[o++] = (t << 32) | (u64)a;
        atomicAdd(&tile_cnt[t], 1);
        if (r.fmask) tile_faces[t] = 1;       // this tile needs the boundary-face path of the assemble kernel
      }
}

// ------------------------------------------------------------------------------------------------ project (hot, FP64)
// Phase 1 of the iso-surface projection once per crossing element: xi0 = Newton projection of xi = 0 onto {rho = rho_t} inside the
// element (it involves the density field only, so one table serves both element kinds); w = 1 when it converged, else the points of
// the element take the edge fallback of proj_init.  Same arithmetic as the per-lane phase 1 (iso::proj_init_element).
template <int MODE>
__global__ void k_phase1(i64 nact, const ActRec *__restrict__ rec, const int *__restrict__ IEN, const double *__restrict__ rn, double rho_t, double4 *__restrict__ p1) {
  const i64 a = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (a >= nact) return;
  const ActRec r = rec[a];
  if (r.cls != 2) { p1[a] = make_double4(0.0, 0.0, 0.0, 0.0); return; }
  double re[8];
#pragma unroll
  for (int k = 0; k < 8; k++) re[k] = rn[IEN[8 * (i64)r.el + k]];
  iso::HexBox B;
  iso::monomial8(re, B.R);
#pragma unroll
  for (int d = 0; d < 3; d++) { B.c[d] = 0.0; B.h[d] = 1.0; B.hh[d] = 2.0; }
  double gs = fabs(rho_t);
#pragma unroll
  for (int k = 0; k < 8; k++) gs = fmax(gs, fabs(re[k]));
  gs = fmax(gs, 1.0);
  iso::ProjState S;
  const bool ok = iso::proj_init_element<iso::HexBox, MODE>(B, rho_t, gs, S);
  p1[a] = make_double4(S.xi[0], S.xi[1], S.xi[2], ok ? 1.0 : 0.0);
}
// one warp per 32-point chunk of a crossing element.  BOX: the variant for axis-aligned box elements (iso::HexBox: 17 element
// constants instead of 32, about half the FP64 work per iteration); kind_check != 0: the mesh holds both kinds of elements and
// each of the two launches leaves the chunks of the other kind alone (ebox[e] = 1 for boxes, built with the mesh).
// MODE: solver variant (r2s_iso.cuh) -- bit 0 FAST restoration (no confirming evaluation, same results), bit 1 one code path for all
// tangent-step cases (results equal to rounding), bit 2 scaled box form (HexBox only), bit 3 the distance of a point in a tile without
// boundary-face elements goes straight into dist[] by a 64-bit atomicMin (the minimum over the pairs is order independent there; the
// pair buffer and the replay of k_assemble are then needed for the tiles with boundary faces only -- as in the lane-refill kernel).
// P1: phase 1 of the solver (Newton projection of xi = 0 onto the iso-surface, the same for every point of an element)
// comes from the per-element table built by k_phase1 instead of being recomputed by every lane of every chunk.
template <bool WANT_XP, int MINB, bool SMEM_A, bool BOX, int MODE, bool P1>
__global__ void __launch_bounds__(128, MINB) k_project_hex8(i64 nitems, i64 nact, const ActRec *__restrict__ rec, const i64 *__restrict__ choff,
                                                      const int *__restrict__ IEN, const double *__restrict__ X, const double *__restrict__ rn, GridDev g,
                                                      double rho_t, double *__restrict__ pairbuf, double *__restrict__ pairxp, u64 *__restrict__ counters,
                                                      const unsigned char *__restrict__ ebox, int kind_check, const double4 *__restrict__ p1,
                                                      const unsigned char *__restrict__ tile_faces, double *__restrict__ dist) {
  i64 item = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (item >= nitems) return;
  // binary search: last a with choff[a] <= item
  i64 lo = 0, hi = nact - 1;
  while (lo < hi) { i64 mid = (lo + hi + 1) >> 1; if (choff[mid] <= item) lo = mid; else hi = mid - 1; }
  const ActRec r = rec[lo];
  if (kind_check && (ebox[r.el] != 0) != BOX) return;
  int chunk = (int)(item - choff[lo]);
  // element data: lane l < 8 loads node l; monomial coefficients assembled through shuffles
  double v[4] = {0, 0, 0, 0};
  if (lane < 8) { i64 n = IEN[8 * (i64)r.el + lane]; v[0] = X[3 * n]; v[1] = X[3 * n + 1]; v[2] = X[3 * n + 2]; v[3] = rn[n]; }
  // MODE bit 1 (the variants that are equal to the default only up to rounding anyway): element-local coordinates -- node 0 is subtracted
  // from the nodes and from the grid point (exact in floating point for any reasonable mesh), so X(xi) - x no longer cancels the mesh
  // offset (tests/test_iso_host.py::test_coordinate_offset_sensitivity)
  double org[3] = {0, 0, 0};
  if constexpr ((MODE & 2) != 0 && !WANT_XP) {
#pragma unroll
    for (int d = 0; d < 3; d++) { org[d] = __shfl_sync(0xffffffffu, v[d], 0); v[d] -= org[d]; }
  }
  // SMEM_A: the warp-uniform monomial coefficients live in shared memory (broadcast reads) instead of 64 registers per thread
  __shared__ double sA[SMEM_A ? 4 : 1][4][8];
  double Ar[SMEM_A ? 1 : 4][8], re[8];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    double nv[8], Ac[8];
#pragma unroll
    for (int k = 0; k < 8; k++) nv[k] = __shfl_sync(0xffffffffu, v[c], k);
    iso::monomial8(nv, Ac);
    if (SMEM_A) { if (lane < 8) sA[threadIdx.x >> 5][c][lane] = Ac[lane & 7]; }
    else {
#pragma unroll
      for (int k = 0; k < 8; k++) Ar[SMEM_A ? 0 : c][k] = Ac[k];
    }
    if (c == 3) {
#pragma unroll
      for (int k = 0; k < 8; k++) re[k] = nv[k];
    }
  }
  if (SMEM_A) __syncwarp();
  const double (*A)[8] = SMEM_A ? (const double (*)[8])sA[threadIdx.x >> 5] : (const double (*)[8])Ar;
  double gs = fabs(rho_t);
#pragma unroll
  for (int k = 0; k < 8; k++) gs = fmax(gs, fabs(re[k]));
  gs = fmax(gs, 1.0);
  int nx = r.pe[0] - r.ps[0], ny = r.pe[1] - r.ps[1], nz = r.pe[2] - r.ps[2];
  i64 vol = (i64)nx * ny * nz, li = (i64)chunk * 32 + lane;
  int nit = 0; bool okc = true;
  if (li < vol) {
    int i = (int)(li % nx), j = (int)((li / nx) % ny), k = (int)(li / ((i64)nx * ny));
    double x[3] = {g.pc[g.pc_off[0] + r.ps[0] + i], g.pc[g.pc_off[1] + r.ps[1] + j], g.pc[g.pc_off[2] + r.ps[2] + k]};
    if constexpr ((MODE & 2) != 0 && !WANT_XP) { x[0] -= org[0]; x[1] -= org[1]; x[2] -= org[2]; }
    double xi[3], p[3] = {0, 0, 0};
    double xi0[3] = {0, 0, 0}; bool ok0 = false;
    if (P1) { const double4 q = p1[lo]; xi0[0] = q.x; xi0[1] = q.y; xi0[2] = q.z; ok0 = q.w != 0.0; }
    if constexpr (BOX && (MODE & 4) != 0) {      // scaled box form (no closest point: the driver never combines it with WANT_XP)
      iso::HexBoxS B; iso::make_box_scaled(A, x, B);
      if (P1) okc = iso::project_hex8_from<iso::HexBoxS, MODE>(B, re, c_hex_sg, c_hex_edges, x, rho_t, gs, xi0, ok0, xi, nit);
      else okc = iso::project_hex8<iso::HexBoxS, MODE>(B, re, c_hex_sg, c_hex_edges, x, rho_t, gs, xi, nit);
      const double dd = sqrt(iso::eval_f(B, x, xi));
      bool direct = false;
      if constexpr ((MODE & 8) != 0) {
        const int pi0 = r.ps[0] + i, pi1 = r.ps[1] + j, pi2 = r.ps[2] + k;
        direct = tile_faces[((i64)(pi2 / TILE_Z) * g.nt[1] + pi1 / TILE_Y) * g.nt[0] + pi0 / TILE_X] == 0;
        if (direct) atomicMin((u64 *)&dist[((i64)pi2 * g.np[1] + pi1) * g.np[0] + pi0], (u64)__double_as_longlong(dd));
      }
      if (!direct) pairbuf[r.pair_off + li] = dd;
    } else {
    if (BOX) {
      iso::HexBox B; iso::make_box(A, B);
      if (P1) okc = iso::project_hex8_from<iso::HexBox, MODE>(B, re, c_hex_sg, c_hex_edges, x, rho_t, gs, xi0, ok0, xi, nit);
      else okc = iso::project_hex8<iso::HexBox, MODE>(B, re, c_hex_sg, c_hex_edges, x, rho_t, gs, xi, nit);
      iso::eval_pos(B, xi, p);
    } else {
      const iso::HexTri T{A};
      if (P1) okc = iso::project_hex8_from<iso::HexTri, MODE>(T, re, c_hex_sg, c_hex_edges, x, rho_t, gs, xi0, ok0, xi, nit);
      else okc = iso::project_hex8<iso::HexTri, MODE>(T, re, c_hex_sg, c_hex_edges, x, rho_t, gs, xi, nit);
      iso::eval_pos(T, xi, p);
    }
    double d0 = x[0] - p[0], d1 = x[1] - p[1], d2 = x[2] - p[2];
    const double dd = sqrt(fma(d2, d2, fma(d1, d1, d0 * d0)));
    bool direct = false;
    if constexpr ((MODE & 8) != 0 && !WANT_XP) {
      const int pi0 = r.ps[0] + i, pi1 = r.ps[1] + j, pi2 = r.ps[2] + k;
      direct = tile_faces[((i64)(pi2 / TILE_Z) * g.nt[1] + pi1 / TILE_Y) * g.nt[0] + pi0 / TILE_X] == 0;
      if (direct) atomicMin((u64 *)&dist[((i64)pi2 * g.np[1] + pi1) * g.np[0] + pi0], (u64)__double_as_longlong(dd));
    }
    if (!direct) pairbuf[r.pair_off + li] = dd;
    if (WANT_XP) { pairxp[3 * (r.pair_off + li)] = p[0]; pairxp[3 * (r.pair_off + li) + 1] = p[1]; pairxp[3 * (r.pair_off + li) + 2] = p[2]; }
    }
  }
  // statistics: iterations and failures (one atomic per warp)
  int its = nit > 0 ? nit : 0;
  for (int o = 16; o > 0; o >>= 1) its += __shfl_down_sync(0xffffffffu, its, o);
  unsigned bad = __ballot_sync(0xffffffffu, !okc);
  if (lane == 0) { u64 *cs = counters + 8 * (1 + (blockIdx.x & 1023)); atomicAdd(&cs[2], (u64)its); if (bad) atomicAdd(&cs[3], (u64)__popc(bad)); }      // statistics spread over 1024 slots
}
// Lane-refill variant (used when the closest points xp are not requested): ONE WARP PER CROSSING ELEMENT.  The warp keeps
// the element's monomial coefficients in registers and its lanes work through the element's candidate points independently:
// a lane that finishes its point immediately takes the next one, so the lanes of a warp sit at different iterations of
// different points and nobody waits for the slowest projection of a 32-point chunk.  Points inside the element's AABB are
// taken first; for the others the distance to the AABB is a lower bound of the result, and a pair whose bound already
// exceeds the voxel's current minimum cannot lower it and is skipped (exact: the result of evalDistances is the minimum over
// the pairs).  The running minimum per voxel is a 64-bit atomicMin on the bit pattern of the non-negative double.  Voxels of
// tiles that hold boundary-face elements keep the pair-buffer path (their replay is order dependent, see k_assemble).
__global__ void k_fill_f64(i64 n, double *__restrict__ a, double v) {
  i64 i = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (i < n) a[i] = v;
}
template <int MINB, bool BOX, int MODE>
__global__ void __launch_bounds__(128, MINB) k_project_hex8_min(i64 nact, const ActRec *__restrict__ rec, const int *__restrict__ IEN, const double *__restrict__ X,
                                                          const double *__restrict__ rn, GridDev g, double rho_t, const unsigned char *__restrict__ tile_faces,
                                                          double *__restrict__ pairbuf, double *__restrict__ dist, u64 *__restrict__ counters,
                                                          const unsigned char *__restrict__ ebox, int kind_check) {
  const i64 a = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
  if (a >= nact) return;
  const ActRec r = rec[a];
  if (r.cls != 2) return;
  if (kind_check && (ebox[r.el] != 0) != BOX) return;
  double v[4] = {0, 0, 0, 0};
  if (lane < 8) { i64 n = IEN[8 * (i64)r.el + lane]; v[0] = X[3 * n]; v[1] = X[3 * n + 1]; v[2] = X[3 * n + 2]; v[3] = rn[n]; }
  double A[4][8], re[8], lo[3], hi[3];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    double nv[8];
#pragma unroll
    for (int k = 0; k < 8; k++) nv[k] = __shfl_sync(0xffffffffu, v[c], k);
    iso::monomial8(nv, A[c]);
    if (c == 3) {
#pragma unroll
      for (int k = 0; k < 8; k++) re[k] = nv[k];
    } else {
      double l = nv[0], h = nv[0];
#pragma unroll
      for (int k = 1; k < 8; k++) { l = fmin(l, nv[k]); h = fmax(h, nv[k]); }
      lo[c] = l; hi[c] = h;
    }
  }
  double gs = fabs(rho_t);
#pragma unroll
  for (int k = 0; k < 8; k++) gs = fmax(gs, fabs(re[k]));
  gs = fmax(gs, 1.0);
  const double margin0 = 1e-12 * fmax(fmax(fmax(fabs(lo[0]), fabs(hi[0])), fmax(fabs(lo[1]), fabs(hi[1]))), fmax(fabs(lo[2]), fabs(hi[2])));
  const int nx = r.pe[0] - r.ps[0], ny = r.pe[1] - r.ps[1], nz = r.pe[2] - r.ps[2];
  const int vol = nx * ny * nz;
  // element in the solver's form: box elements keep 17 constants, the 32 monomial coefficients are dead after this point
  typedef typename std::conditional<BOX, iso::HexBox, iso::HexTri>::type ElemT;
  ElemT EL;
  if constexpr (BOX) iso::make_box(A, EL); else EL.A = A;
  // phase 1 once per element (it does not depend on the grid point)
  iso::ProjState S0; const bool ok0 = iso::proj_init_element<ElemT, MODE>(EL, rho_t, gs, S0);
  bool busy = false, to_buf = false; iso::ProjState S = S0; double x[3] = {0, 0, 0}; int li = 0; i64 vox = 0;
  int sweep = 0, next = 0, its = 0, nbad = 0, npruned = 0;
  while (true) {
    // ---- refill idle lanes with the next candidate points of the current sweep (sweep 0: points inside the AABB, 1: the rest)
    while (sweep < 2) {
      const unsigned idle = __ballot_sync(0xffffffffu, !busy);
      if (!idle) break;
      const int cand = next + __popc(idle & ((1u << lane) - 1));
      if (!busy && cand < vol) {
        const int i = cand % nx, j = (cand / nx) % ny, k = cand / (nx * ny);
        const int pi0 = r.ps[0] + i, pi1 = r.ps[1] + j, pi2 = r.ps[2] + k;
        const double x0 = g.pc[g.pc_off[0] + pi0], x1 = g.pc[g.pc_off[1] + pi1], x2 = g.pc[g.pc_off[2] + pi2];
        const double d0 = fmax(fmax(lo[0] - x0, x0 - hi[0]), 0.0), d1 = fmax(fmax(lo[1] - x1, x1 - hi[1]), 0.0), d2 = fmax(fmax(lo[2] - x2, x2 - hi[2]), 0.0);
        const double lb2 = fma(d2, d2, fma(d1, d1, d0 * d0));
        if ((sweep == 0) == (lb2 == 0.0)) {
          const i64 vx = ((i64)pi2 * g.np[1] + pi1) * g.np[0] + pi0;
          const i64 t = ((i64)(pi2 / TILE_Z) * g.nt[1] + pi1 / TILE_Y) * g.nt[0] + pi0 / TILE_X;
          const bool tb = tile_faces[t] != 0;
          bool prune = false;
          if (!tb && sweep == 1) { const double cur = dist[vx] * (1.0 + 1e-12) + margin0; prune = lb2 > cur * cur; }      // lower bound already above the voxel's minimum
          if (prune) npruned++;
          else {
            li = cand; busy = true; vox = vx; to_buf = tb; x[0] = x0; x[1] = x1; x[2] = x2; S = S0;
            if (!ok0 && !iso::proj_init<ElemT, MODE>(EL, re, c_hex_sg, c_hex_edges, x, rho_t, gs, S)) { S.f = iso::eval_f(EL, x, S.xi); S.it = 1000; }   // no iso point: xi = 0 is used
          }
        }
      }
      next += __popc(idle);
      if (next >= vol) { sweep++; next = 0; }
    }
    if (!__any_sync(0xffffffffu, busy)) break;
    if (busy) {
      int status = 2;
      if (S.it < 100) status = iso::proj_iter<ElemT, MODE>(EL, x, rho_t, gs, S);
      if (status != 0 || S.it >= 100) {
        if (S.it >= 1000) nbad++; else { its += S.it; if (status != 1) nbad++; }
        const double dd = sqrt(S.f);
        if (to_buf) pairbuf[r.pair_off + li] = dd;
        else atomicMin((u64 *)&dist[vox], (u64)__double_as_longlong(dd));
        busy = false;
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) { its += __shfl_down_sync(0xffffffffu, its, o); nbad += __shfl_down_sync(0xffffffffu, nbad, o); npruned += __shfl_down_sync(0xffffffffu, npruned, o); }
  if (lane == 0) { u64 *cs = counters + 8 * (1 + (blockIdx.x & 1023)); atomicAdd(&cs[2], (u64)its); if (nbad) atomicAdd(&cs[3], (u64)nbad); if (npruned) atomicAdd(&cs[4], (u64)npruned); }
}
template <bool WANT_XP>
__global__ void __launch_bounds__(128) k_project_tet4(i64 nitems, i64 nact, const ActRec *__restrict__ rec, const i64 *__restrict__ choff,
                                                      const int *__restrict__ IEN, const double *__restrict__ X, const double *__restrict__ rn, GridDev g,
                                                      double rho_t, double *__restrict__ pairbuf, double *__restrict__ pairxp, u64 *__restrict__ counters) {
  i64 item = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  if (item >= nitems) return;
  i64 lo = 0, hi = nact - 1;
  while (lo < hi) { i64 mid = (lo + hi + 1) >> 1; if (choff[mid] <= item) lo = mid; else hi = mid - 1; }
  const ActRec r = rec[lo];
  int chunk = (int)(item - choff[lo]);
  double Xe[3][4], re[4];
  for (int a = 0; a < 4; a++) { i64 n = IEN[4 * (i64)r.el + a]; re[a] = rn[n]; for (int d = 0; d < 3; d++) Xe[d][a] = X[3 * n + d]; }
  int nx = r.pe[0] - r.ps[0], ny = r.pe[1] - r.ps[1], nz = r.pe[2] - r.ps[2];
  i64 vol = (i64)nx * ny * nz, li = (i64)chunk * 32 + lane;
  if (li < vol) {
    int i = (int)(li % nx), j = (int)((li / nx) % ny), k = (int)(li / ((i64)nx * ny));
    double x[3] = {g.pc[g.pc_off[0] + r.ps[0] + i], g.pc[g.pc_off[1] + r.ps[1] + j], g.pc[g.pc_off[2] + r.ps[2] + k]}, p[3];
    bool ok = iso::project_tet4(Xe, re, c_tet_isn, x, rho_t, p);
    double dist = -1.0;     // negative = "no projection" (the reference would keep the running value)
    if (ok) { double d0 = x[0] - p[0], d1 = x[1] - p[1], d2 = x[2] - p[2]; dist = sqrt((d0 * d0 + d1 * d1) + d2 * d2); }
    else atomicAdd(&counters[3], 1ull);
    pairbuf[r.pair_off + li] = dist;
    if (WANT_XP) { pairxp[3 * (r.pair_off + li)] = p[0]; pairxp[3 * (r.pair_off + li) + 1] = p[1]; pairxp[3 * (r.pair_off + li) + 2] = p[2]; }
  }
}

// ------------------------------------------------------------------------------------------------ assemble (exact)
struct VoxState { double c; double xp[3]; };
template <bool WANT_XP>
__device__ __forceinline__ void write_value(VoxState &s, double dt, const double xp[3]) {   // WriteValue :44-57
  if (fabs(dt) < fabs(s.c)) { s.c = dt; if (WANT_XP) { s.xp[0] = xp[0]; s.xp[1] = xp[1]; s.xp[2] = xp[2]; } }
}
// IsProjectedOnFullSegment (:78-119)
template <bool WANT_XP, int NEN>
__device__ inline bool projected_on_full_segment(const double Xe[3][NEN], const double re[NEN], double rho_t, const double xp[3], const double x[3], VoxState &s,
                                                 const ex::AffineInv *pre = nullptr) {
  double rho;
  if (NEN == 8) {
    double xi[3], N[8];
    if (pre && pre->affine) ex::affine_inverse_apply(*pre, xp, xi);      // same arithmetic as the general path on an affine element
    else ex::inverse_map_hex8((const double(*)[8])Xe, xp, xi);
    if (!(ex::max3abs(xi[0], xi[1], xi[2]) < 1.001)) return false;
    ex::hex8_shape(xi, N);
    rho = ex::dot8(N, re);
  } else {
    double lc[3];
    if (!ex::inverse_map_tet4((const double(*)[4])Xe, xp, lc)) return false;
    if (!(lc[0] >= 0 && lc[1] >= 0 && lc[2] >= 0 && ex::add(ex::add(lc[0], lc[1]), lc[2]) <= 1.0)) return false;
    double l4 = ex::sub(1.0, ex::add(ex::add(lc[0], lc[1]), lc[2]));
    rho = ex::add(ex::add(ex::add(ex::mul(lc[0], re[0]), ex::mul(lc[1], re[1])), ex::mul(lc[2], re[2])), ex::mul(l4, re[3]));
  }
  if (rho >= rho_t) { write_value<WANT_XP>(s, ex::norm3(ex::sub(x[0], xp[0]), ex::sub(x[1], xp[1]), ex::sub(x[2], xp[2])), xp); return true; }
  return false;
}
// process_triangle_projection! (:628-815) for one grid point
template <bool WANT_XP, int NEN>
__device__ inline void triangle_point(const double Xe[3][NEN], const double re[NEN], double rho_t, bool solid, const double Xt[3][3], const double Et[3][3],
                                      const double n[3], const double x[3], VoxState &s, const ex::AffineInv *pre = nullptr) {
  double lam[3]; ex::barycentric(Xt[0], Xt[1], Xt[2], n, x, lam);
  double xp[3]; bool ok = false;
  double lmin = lam[0]; if (lam[1] < lmin) lmin = lam[1]; if (lam[2] < lmin) lmin = lam[2];
  if (lmin >= 0.0) {
#pragma unroll
    for (int d = 0; d < 3; d++) xp[d] = ex::add(ex::add(ex::mul(lam[0], Xt[0][d]), ex::mul(lam[1], Xt[1][d])), ex::mul(lam[2], Xt[2][d]));
    double dt = ex::norm3(ex::sub(x[0], xp[0]), ex::sub(x[1], xp[1]), ex::sub(x[2], xp[2]));
    if (solid) { if (fabs(dt) < fabs(s.c)) { write_value<WANT_XP>(s, dt, xp); ok = true; } }
    else ok = projected_on_full_segment<WANT_XP, NEN>(Xe, re, rho_t, xp, x, s, pre);
  } else {
#pragma unroll
    for (int j = 0; j < 3; j++) {
      if (ok) continue;      // "break" on the first successful edge
      double L = ex::norm3(Et[j][0], Et[j][1], Et[j][2]);
      double u[3] = {ex::dvd(Et[j][0], L), ex::dvd(Et[j][1], L), ex::dvd(Et[j][2], L)};
      double P = ex::add(ex::add(ex::mul(ex::sub(x[0], Xt[j][0]), u[0]), ex::mul(ex::sub(x[1], Xt[j][1]), u[1])), ex::mul(ex::sub(x[2], Xt[j][2]), u[2]));
      if (P >= 0 && P <= L) {
#pragma unroll
        for (int d = 0; d < 3; d++) xp[d] = ex::add(Xt[j][d], ex::mul(u[d], P));
        double dt = ex::norm3(ex::sub(x[0], xp[0]), ex::sub(x[1], xp[1]), ex::sub(x[2], xp[2]));
        if (solid) { if (fabs(dt) < fabs(s.c)) { write_value<WANT_XP>(s, dt, xp); ok = true; } }
        else ok = projected_on_full_segment<WANT_XP, NEN>(Xe, re, rho_t, xp, x, s, pre);
      }
    }
  }
  if (!ok) {
    double dd[3];
#pragma unroll
    for (int j = 0; j < 3; j++) dd[j] = ex::norm3(ex::sub(x[0], Xt[j][0]), ex::sub(x[1], Xt[j][1]), ex::sub(x[2], Xt[j][2]));
    int idx = 0; if (dd[1] < dd[idx]) idx = 1; if (dd[2] < dd[idx]) idx = 2;
    double dmin = idx == 0 ? dd[0] : (idx == 1 ? dd[1] : dd[2]);
#pragma unroll
    for (int d = 0; d < 3; d++) xp[d] = idx == 0 ? Xt[0][d] : (idx == 1 ? Xt[1][d] : Xt[2][d]);
    if (solid) write_value<WANT_XP>(s, dmin, xp);
    else projected_on_full_segment<WANT_XP, NEN>(Xe, re, rho_t, xp, x, s, pre);
  }
}
// Boundary-face triangles (process_boundary_faces! :489-558): every boundary face of an active element is split into nsn
// triangles around its centroid (:520-529).  The triangles and the grid-point range of their (AABB +- delta) cell range
// (Grid.jl:122-154) do not depend on the grid point, so they are built ONCE per call into a table (one thread per active
// element) instead of once per (grid point, candidate element) inside the assemble kernel.
struct TriRec {
  int ps[3], pe[3];      // candidate grid-point range per axis [ps, pe)
  double Xt[3][3];       // vertices (x1, x2, centroid)
  double n[3];           // unit normal
};
__global__ void k_tri_count(i64 nact, const ActRec *__restrict__ rec, int nsn, int *__restrict__ cnt) {
  i64 a = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (a >= nact) return;
  cnt[a] = __popc((unsigned)rec[a].fmask) * nsn;
}
template <int NEN>
__global__ void k_tri_records(i64 nact, ActRec *__restrict__ rec, const int *__restrict__ toff, const int *__restrict__ IEN, const double *__restrict__ X, GridDev g,
                              double delta, TriRec *__restrict__ tri) {
  constexpr int NSN = NEN == 8 ? 4 : 3, NES = NEN == 8 ? 6 : 4;
  i64 a = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (a >= nact) return;
  ActRec r = rec[a];
  rec[a].tri_off = toff[a];
  if (!r.fmask) return;
  double Xe[3][NEN];
  for (int q = 0; q < NEN; q++) { i64 n = IEN[NEN * (i64)r.el + q]; for (int d = 0; d < 3; d++) Xe[d][q] = X[3 * n + d]; }
  int o = toff[a];
  for (int sg = 0; sg < NES; sg++) {
    if (!((r.fmask >> sg) & 1)) continue;
    double Xs[NSN][3], Xc[3];
    for (int q = 0; q < NSN; q++) { int ln = NEN == 8 ? c_hex_isn[sg][q] : c_tet_isn[sg][q]; for (int d = 0; d < 3; d++) Xs[q][d] = Xe[d][ln]; }
    for (int d = 0; d < 3; d++) { double t = Xs[0][d]; for (int q = 1; q < NSN; q++) t = ex::add(t, Xs[q][d]); Xc[d] = ex::dvd(t, (double)NSN); }
    for (int q = 0; q < NSN; q++) {
      int q2 = (q + 1) % NSN; TriRec T; double Et[2][3];
      for (int d = 0; d < 3; d++) { T.Xt[0][d] = Xs[q][d]; T.Xt[1][d] = Xs[q2][d]; T.Xt[2][d] = Xc[d]; }
      bool ok = true;
      for (int d = 0; d < 3; d++) {
        double lo = fmin(T.Xt[0][d], fmin(T.Xt[1][d], T.Xt[2][d])), hi = fmax(T.Xt[0][d], fmax(T.Xt[1][d], T.Xt[2][d])); int I0 = 0, I1 = -1;
        ok = ok && ex::cell_range_axis(lo, hi, delta, g.amin[d], g.amax[d], g.N[d], I0, I1);
        if (ok) { T.ps[d] = g.cstart[g.cs_off[d] + I0]; T.pe[d] = g.cstart[g.cs_off[d] + I1 + 1]; } else { T.ps[d] = 0; T.pe[d] = 0; }
      }
      if (!ok) { for (int d = 0; d < 3; d++) { T.ps[d] = 0; T.pe[d] = 0; } }
      for (int d = 0; d < 3; d++) { Et[0][d] = ex::sub(T.Xt[1][d], T.Xt[0][d]); Et[1][d] = ex::sub(T.Xt[2][d], T.Xt[1][d]); }
      T.n[0] = ex::sub(ex::mul(Et[0][1], Et[1][2]), ex::mul(Et[0][2], Et[1][1]));
      T.n[1] = ex::sub(ex::mul(Et[0][2], Et[1][0]), ex::mul(Et[0][0], Et[1][2]));
      T.n[2] = ex::sub(ex::mul(Et[0][0], Et[1][1]), ex::mul(Et[0][1], Et[1][0]));
      double nn = ex::norm3(T.n[0], T.n[1], T.n[2]);
      T.n[0] = ex::dvd(T.n[0], nn); T.n[1] = ex::dvd(T.n[1], nn); T.n[2] = ex::dvd(T.n[2], nn);
      tri[o++] = T;
    }
  }
}
// process_boundary_faces! for one grid point (point index pi[3], coordinate x): walks the element's triangle records
template <bool WANT_XP, int NEN>
__device__ inline void boundary_faces_point(const ActRec &r, const TriRec *__restrict__ tri, const int *__restrict__ IEN, const double *__restrict__ X,
                                            const double *__restrict__ rn, double rho_t, const int pi[3], const double x[3], VoxState &s) {
  constexpr int NSN = NEN == 8 ? 4 : 3;
  const int ntri = __popc((unsigned)r.fmask) * NSN;
  bool loaded = false;
  double Xe[3][NEN], re[NEN];
  for (int t = 0; t < ntri; t++) {
    const TriRec &T = tri[r.tri_off + t];
    if (pi[0] < T.ps[0] || pi[0] >= T.pe[0] || pi[1] < T.ps[1] || pi[1] >= T.pe[1] || pi[2] < T.ps[2] || pi[2] >= T.pe[2]) continue;
    if (r.cls != 1 && !loaded) {      // the element itself is only needed for the rho-test of crossing elements (:92-113)
      for (int a = 0; a < NEN; a++) { i64 n = IEN[NEN * (i64)r.el + a]; re[a] = rn[n]; for (int d = 0; d < 3; d++) Xe[d][a] = X[3 * n + d]; }
      loaded = true;
    }
    double Xt[3][3], Et[3][3], n[3];
    for (int d = 0; d < 3; d++) { Xt[0][d] = T.Xt[0][d]; Xt[1][d] = T.Xt[1][d]; Xt[2][d] = T.Xt[2][d]; n[d] = T.n[d]; }
    if (r.cls == 1) {
      // Solid element: every candidate of this triangle (face, edge or vertex projection) only replaces the running value if it is
      // strictly smaller, and each of them is at least the distance from x to the triangle's bounding box.  If that bound is not
      // below the running value the triangle cannot change anything (exact; the margin covers the rounding of the candidates).
      double lb2 = 0.0;
#pragma unroll
      for (int d = 0; d < 3; d++) {
        const double lo = fmin(Xt[0][d], fmin(Xt[1][d], Xt[2][d])), hi = fmax(Xt[0][d], fmax(Xt[1][d], Xt[2][d]));
        const double e = fmax(fmax(lo - x[d], x[d] - hi), 0.0);
        lb2 = fma(e, e, lb2);
      }
      const double cur = fabs(s.c) * (1.0 + 1e-12);
      if (lb2 * (1.0 - 1e-12) > cur * cur) continue;
    }
    for (int d = 0; d < 3; d++) { Et[0][d] = ex::sub(Xt[1][d], Xt[0][d]); Et[1][d] = ex::sub(Xt[2][d], Xt[1][d]); Et[2][d] = ex::sub(Xt[0][d], Xt[2][d]); }
    triangle_point<WANT_XP, NEN>(Xe, re, rho_t, r.cls == 1, Xt, Et, n, x, s);
  }
}

// Boundary faces of CROSSING elements (process_boundary_faces!(..., false), :584).  For a crossing element every candidate of
// process_triangle_projection! is accepted or rejected by the rho-test alone (IsProjectedOnFullSegment, :78-119) -- never by
// the running minimum -- and accepted candidates only ever lower the minimum.  So the element's face contribution to a grid
// point is ONE number, min over its boundary triangles, independent of the order of anything else, and it can be folded into
// the element's entry of the pair buffer: pair = min(iso distance, face candidates).  One warp per crossing element with
// boundary faces, lanes over the element's candidate points; the order-dependent replay (k_assemble) is then left with the
// faces of SOLID elements only, which need no inverse map.
template <int NEN>
__global__ void __launch_bounds__(128) k_faces_crossing(i64 nact, const ActRec *__restrict__ rec, const TriRec *__restrict__ tri, const int *__restrict__ IEN,
                                                        const double *__restrict__ X, const double *__restrict__ rn, GridDev g, double rho_t, double *__restrict__ pairbuf) {
  constexpr int NSN = NEN == 8 ? 4 : 3;
  const i64 a = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5; const int lane = threadIdx.x & 31;
  if (a >= nact) return;
  const ActRec r = rec[a];
  if (r.cls != 2 || !r.fmask) return;
  double Xe[3][NEN], re[NEN];
  for (int q = 0; q < NEN; q++) { i64 n = IEN[NEN * (i64)r.el + q]; re[q] = rn[n]; for (int d = 0; d < 3; d++) Xe[d][q] = X[3 * n + d]; }
  ex::AffineInv pre; pre.affine = 0;
  if (NEN == 8) {      // element-only part of the inverse map, once per warp instead of once per candidate
    double A[3][8];
#pragma unroll
    for (int d = 0; d < 3; d++) ex::mono8((const double *)Xe[d], A[d]);
    ex::affine_inverse_prepare(A, pre);
  }
  const int ntri = __popc((unsigned)r.fmask) * NSN;
  const int nx = r.pe[0] - r.ps[0], ny = r.pe[1] - r.ps[1], nz = r.pe[2] - r.ps[2], vol = nx * ny * nz;
  for (int li = lane; li < vol; li += 32) {
    const int pi[3] = {r.ps[0] + li % nx, r.ps[1] + (li / nx) % ny, r.ps[2] + li / (nx * ny)};
    const double x[3] = {g.pc[g.pc_off[0] + pi[0]], g.pc[g.pc_off[1] + pi[1]], g.pc[g.pc_off[2] + pi[2]]};
    VoxState s; s.c = -R2S_BIG; s.xp[0] = s.xp[1] = s.xp[2] = 0.0;
    for (int t = 0; t < ntri; t++) {
      const TriRec &T = tri[r.tri_off + t];
      if (pi[0] < T.ps[0] || pi[0] >= T.pe[0] || pi[1] < T.ps[1] || pi[1] >= T.pe[1] || pi[2] < T.ps[2] || pi[2] >= T.pe[2]) continue;
      double Xt[3][3], Et[3][3], n[3];
      for (int d = 0; d < 3; d++) { Xt[0][d] = T.Xt[0][d]; Xt[1][d] = T.Xt[1][d]; Xt[2][d] = T.Xt[2][d]; n[d] = T.n[d]; }
      for (int d = 0; d < 3; d++) { Et[0][d] = ex::sub(Xt[1][d], Xt[0][d]); Et[1][d] = ex::sub(Xt[2][d], Xt[1][d]); Et[2][d] = ex::sub(Xt[0][d], Xt[2][d]); }
      triangle_point<false, NEN>(Xe, re, rho_t, false, Xt, Et, n, x, s, &pre);
    }
    if (s.c != -R2S_BIG) {
      const i64 idx = r.pair_off + li;
      const double cur = pairbuf[idx];
      if (!(cur >= 0.0 && cur <= s.c)) pairbuf[idx] = s.c;
    }
  }
}

// One CTA per tile, one thread per grid point; the tile's element list is culled per warp (footprint 8x4x1 points) into
// shared memory in list order, then every lane replays ITS candidates in ascending element order.
#define ACULL_CAP 96
// FACES = false: tiles whose list holds no element with boundary faces (the vast majority) run a light variant with a
// small register footprint; FACES = true handles the others.  Both walk all tiles and skip those of the other kind.
template <bool WANT_XP, int NEN, bool FACES>
__global__ void __launch_bounds__(TILE_VOX) k_assemble(GridDev g, int kz0, int kz1, const unsigned char *__restrict__ tile_faces, const int *__restrict__ tile_ptr, const u64 *__restrict__ keys,
                                                       const ActRec *__restrict__ rec, const TriRec *__restrict__ tri, const int *__restrict__ IEN, const double *__restrict__ X,
                                                       const double *__restrict__ rn, double rho_t, double delta, const double *__restrict__ pairbuf,
                                                       const double *__restrict__ pairxp, double *__restrict__ dist, double *__restrict__ xpo) {
  __shared__ ActRec srec[TILE_VOX / 32][ACULL_CAP];
  const int t = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((tile_faces[t] != 0) != FACES) return;
  const int tx = t % g.nt[0], ty = (t / g.nt[0]) % g.nt[1], tz = t / (g.nt[0] * g.nt[1]);
  const int li = threadIdx.x % TILE_X, lj = (threadIdx.x / TILE_X) % TILE_Y, lk = threadIdx.x / (TILE_X * TILE_Y);
  const int pi[3] = {tx * TILE_X + li, ty * TILE_Y + lj, tz * TILE_Z + lk};
  const bool valid = pi[0] < g.np[0] && pi[1] < g.np[1] && pi[2] < g.np[2] && pi[2] >= kz0 && pi[2] < kz1;
  const int wx0 = tx * TILE_X, wx1 = wx0 + TILE_X - 1, wy0 = ty * TILE_Y + (warp % (TILE_Y / 4)) * 4, wy1 = wy0 + 3, wz = tz * TILE_Z + warp / (TILE_Y / 4);
  double x[3] = {0, 0, 0};
  if (valid) { x[0] = g.pc[g.pc_off[0] + pi[0]]; x[1] = g.pc[g.pc_off[1] + pi[1]]; x[2] = g.pc[g.pc_off[2] + pi[2]]; }
  VoxState s; s.c = -R2S_BIG; s.xp[0] = s.xp[1] = s.xp[2] = 0.0;
  const int p0 = tile_ptr[t], p1 = tile_ptr[t + 1];
  int p = p0;
  while (p < p1) {
    int n = 0;
    while (p < p1 && n <= ACULL_CAP - 32) {
      int idx = p + lane; bool ov = false; ActRec r;
      if (idx < p1) {
        r = rec[(int)(keys[idx] & 0xffffffffull)];
        ov = r.ps[0] <= wx1 && r.pe[0] > wx0 && r.ps[1] <= wy1 && r.pe[1] > wy0 && r.ps[2] <= wz && r.pe[2] > wz;
      }
      unsigned m = __ballot_sync(0xffffffffu, ov);
      if (ov) srec[warp][n + __popc(m & ((1u << lane) - 1))] = r;
      n += __popc(m); p += 32;
    }
    __syncwarp();
    // which culled records contain THIS lane's point: one uniform sweep (broadcast reads)
    unsigned mk0 = 0, mk1 = 0, mk2 = 0;
    for (int q = 0; q < n; q++) {
      const ActRec &r = srec[warp][q];
      if (valid && pi[0] >= r.ps[0] && pi[0] < r.pe[0] && pi[1] >= r.ps[1] && pi[1] < r.pe[1] && pi[2] >= r.ps[2] && pi[2] < r.pe[2]) {
        if (q < 32) mk0 |= 1u << q; else if (q < 64) mk1 |= 1u << (q - 32); else mk2 |= 1u << (q - 64);
      }
    }
    // Every lane walks ITS records in list order (= ascending element index) and, inside a record, its boundary triangles in
    // order, then takes the record's pair-buffer entry -- exactly the reference's sequence for that grid point.  The lanes do
    // not wait for each other's records: in each round every lane brings its own next triangle to triangle_point, so the
    // expensive part runs with (nearly) full warps instead of only the lanes that happen to share the current element.
    constexpr int NSN = NEN == 8 ? 4 : 3;
    int w = 0, pos = -1, t = 0, ntri = 0; unsigned cur = mk0; bool more = valid;
    ActRec r; r.cls = 0; r.tri_off = 0; r.pair_off = 0; r.fmask = 0; r.el = 0;
    double Xe[3][NEN], re[NEN]; bool loaded = false;
    while (true) {
      int ti = -1;                                   // index of this lane's next triangle (work item of this round)
      while (more && ti < 0) {
        if (t < ntri) {
          const TriRec &T = tri[r.tri_off + t]; const int tcur = t; t++;
          if (pi[0] < T.ps[0] || pi[0] >= T.pe[0] || pi[1] < T.ps[1] || pi[1] >= T.pe[1] || pi[2] < T.ps[2] || pi[2] >= T.pe[2]) continue;
          if (r.cls == 1) {
            // solid element: no candidate of this triangle can be below the distance to its bounding box; if that is not
            // below the running value the triangle changes nothing (exact; the margin covers the rounding of the candidates)
            double lb2 = 0.0;
#pragma unroll
            for (int d = 0; d < 3; d++) {
              const double lo = fmin(T.Xt[0][d], fmin(T.Xt[1][d], T.Xt[2][d])), hi = fmax(T.Xt[0][d], fmax(T.Xt[1][d], T.Xt[2][d]));
              const double e = fmax(fmax(lo - x[d], x[d] - hi), 0.0);
              lb2 = fma(e, e, lb2);
            }
            const double cv = fabs(s.c) * (1.0 + 1e-12);
            if (lb2 * (1.0 - 1e-12) > cv * cv) continue;
          }
          ti = tcur;
        } else {
          if (pos >= 0 && r.cls == 2) {              // the record's faces are done: now its iso distance (:617-621)
            i64 idx = r.pair_off + ((i64)(pi[2] - r.ps[2]) * (r.pe[1] - r.ps[1]) + (pi[1] - r.ps[1])) * (r.pe[0] - r.ps[0]) + (pi[0] - r.ps[0]);
            double dt = pairbuf[idx];
            if (dt >= 0.0 && fabs(dt) < fabs(s.c)) {
              s.c = dt;
              if (WANT_XP) { s.xp[0] = pairxp[3 * idx]; s.xp[1] = pairxp[3 * idx + 1]; s.xp[2] = pairxp[3 * idx + 2]; }
            }
          }
          while (w < 3 && cur == 0) { w++; cur = (w == 1) ? mk1 : (w == 2 ? mk2 : 0u); }
          if (w >= 3) { more = false; pos = -1; break; }
          const int bq = __ffs(cur) - 1; cur &= cur - 1; pos = w * 32 + bq;
          r = srec[warp][pos];
          t = 0; loaded = false;
          ntri = (FACES && r.fmask && (WANT_XP || r.cls == 1)) ? __popc((unsigned)r.fmask) * NSN : 0;      // crossing faces are folded into the pair buffer unless xp is wanted
        }
      }
      if (!__any_sync(0xffffffffu, ti >= 0)) break;
      if (ti >= 0) {
        const TriRec &T = tri[r.tri_off + ti];
        if (WANT_XP && r.cls != 1 && !loaded) {      // the element itself is only needed for the rho-test of crossing elements (:92-113), replayed here only when xp is wanted
          for (int a = 0; a < NEN; a++) { i64 nd = IEN[NEN * (i64)r.el + a]; re[a] = rn[nd]; for (int d = 0; d < 3; d++) Xe[d][a] = X[3 * nd + d]; }
          loaded = true;
        }
        double Xt[3][3], Et[3][3], nn[3];
        for (int d = 0; d < 3; d++) { Xt[0][d] = T.Xt[0][d]; Xt[1][d] = T.Xt[1][d]; Xt[2][d] = T.Xt[2][d]; nn[d] = T.n[d]; }
        for (int d = 0; d < 3; d++) { Et[0][d] = ex::sub(Xt[1][d], Xt[0][d]); Et[1][d] = ex::sub(Xt[2][d], Xt[1][d]); Et[2][d] = ex::sub(Xt[0][d], Xt[2][d]); }
        triangle_point<WANT_XP, NEN>(Xe, re, rho_t, WANT_XP ? r.cls == 1 : true, Xt, Et, nn, x, s);
      }
    }
    __syncwarp();
  }
  if (valid) {
    i64 v = ((i64)pi[2] * g.np[1] + pi[1]) * g.np[0] + pi[0];
    dist[v] = fabs(s.c);
    if (WANT_XP) { xpo[3 * v] = s.xp[0]; xpo[3 * v + 1] = s.xp[1]; xpo[3 * v + 2] = s.xp[2]; }
  }
}

// ------------------------------------------------------------------------------------------------ host driver
int r2s_dev_eval_distances(r2s_ctx *ctx, double rho_t, double delta_factor, bool want_xp) {
  if (!ctx->has_grid) FAIL("r2s_set_grid has not been called");
  if (ctx->nel == 0) FAIL("r2s_set_mesh has not been called");
  const GridDev &g = ctx->g; cudaStream_t st = ctx->stream;
  i64 nel = ctx->nel; int nen = ctx->nen;
  double delta = delta_factor * g.cell;
  int kz0 = (int)ctx->k0, kz1 = (int)ctx->k1;
  CK(cudaEventRecord(ctx->ev[0], st));
  CK(ctx->cls.reserve((size_t)nel));
  CK(ctx->act_flag.reserve(sizeof(int) * (size_t)(nel + 1)));
  CK(ctx->act_idx.reserve(sizeof(int) * (size_t)(nel + 1)));
  CK(ctx->counters.reserve(sizeof(u64) * 8 * 1025));      // slot 0: element classes; slots 1..1024: projection statistics
  CK(cudaMemsetAsync(ctx->counters.p, 0, sizeof(u64) * 8 * 1025, st));
  CK(cudaMemsetAsync(ctx->act_flag.as<int>() + nel, 0, sizeof(int), st));
  // z-interval of this rank's planes with a margin of (delta + 2 cells): only elements that can reach them are classified
  double zlo = -1e300, zhi = 1e300;
  if (kz0 > 0 || kz1 < g.np[2]) { const double m = delta + 2.0 * g.cell; zlo = ctx->h_pc[2][(size_t)kz0] - m; zhi = ctx->h_pc[2][(size_t)kz1 - 1] + m; }
  k_classify<<<cdiv(nel, 256), 256, 0, st>>>(nel, nen, ctx->IEN32.as<int>(), ctx->rho_n.as<double>(), ctx->fbnd.as<unsigned char>(), ctx->ezr.as<double2>(), zlo, zhi, rho_t,
                                             ctx->cls.as<unsigned char>(), ctx->act_flag.as<int>(), ctx->counters.as<i64>()); LAUNCH_CHECK();
  if (r2s_scan_exclusive_i32(ctx, ctx->act_flag.as<int>(), ctx->act_idx.as<int>(), nel + 1)) return 1;
  int nact_i = 0;
  CK(cudaMemcpyAsync(&nact_i, ctx->act_idx.as<int>() + nel, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  i64 nact = nact_i;
  CK(ctx->dist.reserve(sizeof(double) * (size_t)g.ngp));
  if (want_xp) CK(ctx->xp.reserve(sizeof(double) * 3 * (size_t)g.ngp));
  CK(ctx->tile_ptr.reserve(sizeof(int) * (size_t)(g.ntiles + 2)));
  CK(cudaMemsetAsync(ctx->tile_ptr.p, 0, sizeof(int) * (size_t)(g.ntiles + 2), st));
  CK(ctx->tile_faces.reserve((size_t)g.ntiles + 16));
  CK(cudaMemsetAsync(ctx->tile_faces.p, 0, (size_t)g.ntiles, st));
  i64 npairs = 0, nkeys = 0, nitems = 0;
  u64 *sorted = nullptr;
  if (nact > 0) {
    CK(ctx->act_rec.reserve(sizeof(ActRec) * (size_t)nact));
    CK(ctx->cnt_a.reserve(sizeof(i64) * 3 * (size_t)(nact + 1)));
    CK(ctx->cnt_b.reserve(sizeof(i64) * 3 * (size_t)(nact + 1)));
    i64 *ntile = ctx->cnt_a.as<i64>(), *npair = ntile + (nact + 1), *nchunk = npair + (nact + 1);
    i64 *toff = ctx->cnt_b.as<i64>(), *poff = toff + (nact + 1), *choff = poff + (nact + 1);
    CK(cudaMemsetAsync(ctx->cnt_a.p, 0, sizeof(i64) * 3 * (size_t)(nact + 1), st));
    k_act_records<<<cdiv(nel, 256), 256, 0, st>>>(nel, nen, ctx->IEN32.as<int>(), ctx->X.as<double>(), ctx->cls.as<unsigned char>(), ctx->fbnd.as<unsigned char>(),
                                                  ctx->act_flag.as<int>(), ctx->act_idx.as<int>(), g, delta, kz0, kz1, ctx->act_rec.as<ActRec>(), ntile, npair, nchunk);
    LAUNCH_CHECK();
    if (r2s_scan_exclusive_i64(ctx, ntile, toff, nact + 1)) return 1;
    if (r2s_scan_exclusive_i64(ctx, npair, poff, nact + 1)) return 1;
    if (r2s_scan_exclusive_i64(ctx, nchunk, choff, nact + 1)) return 1;
    i64 h3[3];
    CK(cudaMemcpyAsync(&h3[0], toff + nact, sizeof(i64), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h3[1], poff + nact, sizeof(i64), cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(&h3[2], choff + nact, sizeof(i64), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    nkeys = h3[0]; npairs = h3[1]; nitems = h3[2];
    if (nkeys >= (1ll << 31) || npairs >= (1ll << 40)) FAIL("distance binning: problem too large for one slab (key/pair count overflow)");
    CK(ctx->keys.reserve(sizeof(u64) * (size_t)(nkeys + 1)));
    CK(ctx->keys_alt.reserve(sizeof(u64) * (size_t)(nkeys + 1)));
    CK(ctx->pairbuf.reserve(sizeof(double) * (size_t)(npairs + 1)));
    if (want_xp) CK(ctx->pairxp.reserve(sizeof(double) * 3 * (size_t)(npairs + 1)));
    k_emit_keys<<<cdiv(nact, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), toff, poff, g, ctx->keys.as<u64>(), ctx->tile_ptr.as<int>() + 1, ctx->tile_faces.as<unsigned char>()); LAUNCH_CHECK();
    int tbits = 1; while ((1ll << tbits) < g.ntiles) tbits++;
    if (nkeys > 0) { if (r2s_sort_keys_u64(ctx, ctx->keys.as<u64>(), ctx->keys_alt.as<u64>(), nkeys, 32 + tbits, &sorted)) return 1; }
    // boundary-face triangle table
    CK(ctx->tri_cnt.reserve(sizeof(int) * 2 * (size_t)(nact + 1)));
    int *tcnt = ctx->tri_cnt.as<int>(), *toff32 = tcnt + (nact + 1);
    CK(cudaMemsetAsync(tcnt + nact, 0, sizeof(int), st));
    k_tri_count<<<cdiv(nact, 256), 256, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), ctx->nsn, tcnt); LAUNCH_CHECK();
    if (r2s_scan_exclusive_i32(ctx, tcnt, toff32, nact + 1)) return 1;
    int ntri = 0;
    CK(cudaMemcpyAsync(&ntri, toff32 + nact, sizeof(int), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(ctx->tri_rec.reserve(sizeof(TriRec) * (size_t)(ntri + 1)));
    if (nen == 8) k_tri_records<8><<<cdiv(nact, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), toff32, ctx->IEN32.as<int>(), ctx->X.as<double>(), g, delta, ctx->tri_rec.as<TriRec>());
    else k_tri_records<4><<<cdiv(nact, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), toff32, ctx->IEN32.as<int>(), ctx->X.as<double>(), g, delta, ctx->tri_rec.as<TriRec>());
    LAUNCH_CHECK();
  }
  // tile_ptr[t+1] currently holds the count of tile t (tile_ptr[0] = 0): inclusive scan in place == exclusive offsets
  {
    DevBuf &tmp = ctx->s_cnt;   // reuse as scratch
    CK(tmp.reserve(sizeof(int) * (size_t)(g.ntiles + 2)));
    if (r2s_scan_exclusive_i32(ctx, ctx->tile_ptr.as<int>() + 1, tmp.as<int>(), g.ntiles + 1)) return 1;
    CK(cudaMemcpyAsync(ctx->tile_ptr.as<int>(), tmp.as<int>(), sizeof(int) * (size_t)(g.ntiles + 1), cudaMemcpyDeviceToDevice, st));
  }
  CK(cudaEventRecord(ctx->ev[1], st));
  // HEX8 without xp: lane-refill projection with per-voxel atomicMin (face-free tiles) + exact replay of the tiles with boundary faces.
  // (The R2S_PROJ* tuning knobs are read on every call so that one process can time the variants side by side, tools/ab_project.py.)
  const bool refill = getenv("R2S_PROJ") && atoi(getenv("R2S_PROJ")) == 1;       // experimental lane-refill projection (not faster yet)
  const bool minpath = (refill && nen == 8 && !want_xp);
  // chunk kernel with direct atomicMin output (opt-in, MODE bit 3): dist[] starts at BIG, k_assemble replays the tiles with boundary faces only
  const bool atom = !minpath && nen == 8 && !want_xp && getenv("R2S_PROJ_ATOM") && atoi(getenv("R2S_PROJ_ATOM")) == 1 && !(getenv("R2S_PROJ_SMEMA") && atoi(getenv("R2S_PROJ_SMEMA")) == 1);
  if (atom) {
    i64 v0 = (i64)kz0 * g.np[0] * g.np[1], nv = (i64)(kz1 - kz0) * g.np[0] * g.np[1];
    k_fill_f64<<<cdiv(nv, 256), 256, 0, st>>>(nv, ctx->dist.as<double>() + v0, R2S_BIG); LAUNCH_CHECK();
  }
  if (minpath) {
    i64 v0 = (i64)kz0 * g.np[0] * g.np[1], nv = (i64)(kz1 - kz0) * g.np[0] * g.np[1];
    k_fill_f64<<<cdiv(nv, 256), 256, 0, st>>>(nv, ctx->dist.as<double>() + v0, R2S_BIG); LAUNCH_CHECK();
    if (nact > 0 && npairs > 0) {
      // occupancy variant (registers per thread 255 / 168 / 128): R2S_PROJ_MINB = 2, 3, 4 (tuning knob, default from measurements)
      const int minbr = getenv("R2S_PROJ_MINB") ? atoi(getenv("R2S_PROJ_MINB")) : 3;
      // box elements (n_box of them, flag built with the mesh) go through the HexBox variant; a mixed mesh takes both launches
      const bool use_box_r = !(getenv("R2S_PROJ_BOX") && atoi(getenv("R2S_PROJ_BOX")) == 0);
      const i64 nbx = use_box_r ? ctx->n_box : 0; const int kc = (nbx > 0 && nbx < nel) ? 1 : 0;
      const bool uni_r = getenv("R2S_PROJ_UNI") && atoi(getenv("R2S_PROJ_UNI")) == 1;
      const bool fast_r = uni_r || (getenv("R2S_PROJ_FAST") && atoi(getenv("R2S_PROJ_FAST")) == 1);
#define PMIN(MB, BX, MD) k_project_hex8_min<MB, BX, MD><<<cdiv(nact * 32, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), ctx->IEN32.as<int>(), ctx->X.as<double>(), ctx->rho_n.as<double>(), g, rho_t, \
                                                              ctx->tile_faces.as<unsigned char>(), ctx->pairbuf.as<double>(), ctx->dist.as<double>(), ctx->counters.as<u64>(), ctx->ebox.as<unsigned char>(), kc)
#define PMINF(MB, BX) do { if (uni_r) PMIN(MB, BX, 3); else if (fast_r) PMIN(MB, BX, 1); else PMIN(MB, BX, 0); } while (0)
      if (nbx < nel) { if (minbr <= 2) PMINF(2, false); else if (minbr == 3) PMINF(3, false); else PMINF(4, false); LAUNCH_CHECK(); }
      if (nbx > 0) { if (minbr <= 2) PMINF(2, true); else if (minbr == 3) PMINF(3, true); else PMINF(4, true); }
#undef PMINF
#undef PMIN
      LAUNCH_CHECK();
    }
  } else if (nitems > 0) {
    i64 *choff = ctx->cnt_b.as<i64>() + 2 * (nact + 1);
    int nb = cdiv(nitems * 32, 128);
#define PROJ(KERN, XP) KERN<XP><<<nb, 128, 0, st>>>(nitems, nact, ctx->act_rec.as<ActRec>(), choff, ctx->IEN32.as<int>(), ctx->X.as<double>(), \
                                                   ctx->rho_n.as<double>(), g, rho_t, ctx->pairbuf.as<double>(), ctx->pairxp.as<double>(), ctx->counters.as<u64>())
    // variants of the HEX8 kernel: R2S_PROJ_MINB = 2..6 CTAs/SM (255 / 168 / 128 / 102 / 85 registers), R2S_PROJ_SMEMA = 1 keeps the
    // element's monomial coefficients in shared memory
    const int minb = getenv("R2S_PROJ_MINB") ? atoi(getenv("R2S_PROJ_MINB")) : 4;      // measured: 4 CTAs/SM is 22% faster than 2
    const bool smema = getenv("R2S_PROJ_SMEMA") && atoi(getenv("R2S_PROJ_SMEMA")) == 1;
#define PROJH6(XP, MB, SA, BX, MD, PP) k_project_hex8<XP, MB, SA, BX, MD, PP><<<nb, 128, 0, st>>>(nitems, nact, ctx->act_rec.as<ActRec>(), choff, ctx->IEN32.as<int>(), ctx->X.as<double>(), \
                                                   ctx->rho_n.as<double>(), g, rho_t, ctx->pairbuf.as<double>(), ctx->pairxp.as<double>(), ctx->counters.as<u64>(), ctx->ebox.as<unsigned char>(), kc, p1tab, ctx->tile_faces.as<unsigned char>(), ctx->dist.as<double>())
#define PROJH(XP, MB, SA, BX, FS) PROJH6(XP, MB, SA, BX, 0, false)
    // the opt-in variants (FAST solver and / or phase 1 from the table), instantiated for the occupancies worth measuring
#define PROJO(MB, BX) do { \
      if (atom && scaled && BX) PROJH6(false, MB, false, BX, 15, true); \
      else if (atom) PROJH6(false, MB, false, BX, 11, true); \
      else if (scaled && BX) { if (use_p1) PROJH6(false, MB, false, BX, 7, true); else PROJH6(false, MB, false, BX, 7, false); } \
      else if (uni) { if (use_p1) PROJH6(false, MB, false, BX, 3, true); else PROJH6(false, MB, false, BX, 3, false); } \
      else if (fast) { if (use_p1) PROJH6(false, MB, false, BX, 1, true); else PROJH6(false, MB, false, BX, 1, false); } \
      else PROJH6(false, MB, false, BX, 0, true); } while (0)
    // Axis-aligned box elements (flag + count built with the mesh) take the HexBox variant of the kernel, the others the general
    // trilinear one; a mesh with both kinds takes both launches, each leaving the other kind's chunks alone.  R2S_PROJ_BOX=0
    // sends everything through the general kernel; R2S_PROJ_BOX_MINB = CTAs/SM of the box variant (4 / 5 / 6).
    // Not yet measured on a GPU, hence opt-in: R2S_PROJ_FAST=1 (FAST restoration), R2S_PROJ_UNI=1 (FAST + one tangent-step code path), R2S_PROJ_SCALED=1 (UNI + the scaled box form HexBoxS), R2S_PROJ_P1=1 (phase 1 from a per-element table).
    const bool use_box = !(getenv("R2S_PROJ_BOX") && atoi(getenv("R2S_PROJ_BOX")) == 0);
    const int minb_box = getenv("R2S_PROJ_BOX_MINB") ? atoi(getenv("R2S_PROJ_BOX_MINB")) : 5;      // measured at n = 256: 92.6 / 86.5 / 89.5 ms for 4 / 5 / 6 CTAs per SM (profiles/r1f_ab_project_variants_n256.jsonl)
    // R2S_PROJ_ATOM=1: UNI + P1 + direct atomicMin output for tiles without boundary-face elements (decided above: `atom`)
    const bool scaled = nen == 8 && !want_xp && !smema && getenv("R2S_PROJ_SCALED") && atoi(getenv("R2S_PROJ_SCALED")) == 1;
    const bool uni = scaled || atom || (nen == 8 && !want_xp && !smema && getenv("R2S_PROJ_UNI") && atoi(getenv("R2S_PROJ_UNI")) == 1);
    const bool fast = uni || (nen == 8 && !want_xp && !smema && getenv("R2S_PROJ_FAST") && atoi(getenv("R2S_PROJ_FAST")) == 1);
    const bool use_p1 = atom || (nen == 8 && !want_xp && !smema && getenv("R2S_PROJ_P1") && atoi(getenv("R2S_PROJ_P1")) == 1);
    const i64 nbx = (use_box && nen == 8) ? ctx->n_box : 0; const int kc = (nbx > 0 && nbx < nel) ? 1 : 0;
    const double4 *p1tab = nullptr;
    if (use_p1) {
      CK(ctx->p1tab.reserve(sizeof(double4) * (size_t)(nact + 1)));
      if (fast) k_phase1<1><<<cdiv(nact, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), ctx->IEN32.as<int>(), ctx->rho_n.as<double>(), rho_t, ctx->p1tab.as<double4>());
      else k_phase1<0><<<cdiv(nact, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), ctx->IEN32.as<int>(), ctx->rho_n.as<double>(), rho_t, ctx->p1tab.as<double4>());
      LAUNCH_CHECK();
      p1tab = ctx->p1tab.as<double4>();
    }
    if (nen == 8) {
      const bool optin = (fast || use_p1) && !want_xp && !smema;
      if (nbx < nel) {
        if (want_xp) PROJH(true, 2, false, false, false);
        else if (smema) { if (minb <= 4) PROJH(false, 4, true, false, false); else if (minb == 5) PROJH(false, 5, true, false, false); else PROJH(false, 6, true, false, false); }
        else if (optin) { if (minb <= 4) PROJO(4, false); else PROJO(5, false); }
        else if (minb <= 2) PROJH(false, 2, false, false, false); else if (minb == 3) PROJH(false, 3, false, false, false); else if (minb == 4) PROJH(false, 4, false, false, false);
        else if (minb == 5) PROJH(false, 5, false, false, false); else PROJH(false, 6, false, false, false);
        if (nbx > 0) LAUNCH_CHECK();
      }
      if (nbx > 0) {
        if (want_xp) PROJH(true, 4, false, true, false);
        else if (optin) { if (minb_box <= 5) PROJO(5, true); else if (minb_box == 6) PROJO(6, true); else PROJO(8, true); }
        else if (minb_box <= 4) PROJH(false, 4, false, true, false); else if (minb_box == 5) PROJH(false, 5, false, true, false); else PROJH(false, 6, false, true, false);
      }
    } else { if (want_xp) PROJ(k_project_tet4, true); else PROJ(k_project_tet4, false); }
    LAUNCH_CHECK();
#undef PROJO
#undef PROJH
#undef PROJH6
#undef PROJ
  }
  if (!want_xp && nact > 0 && npairs > 0) {
    if (nen == 8) k_faces_crossing<8><<<cdiv(nact * 32, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), ctx->tri_rec.as<TriRec>(), ctx->IEN32.as<int>(), ctx->X.as<double>(), ctx->rho_n.as<double>(), g, rho_t, ctx->pairbuf.as<double>());
    else k_faces_crossing<4><<<cdiv(nact * 32, 128), 128, 0, st>>>(nact, ctx->act_rec.as<ActRec>(), ctx->tri_rec.as<TriRec>(), ctx->IEN32.as<int>(), ctx->X.as<double>(), ctx->rho_n.as<double>(), g, rho_t, ctx->pairbuf.as<double>());
    LAUNCH_CHECK();
  }
  CK(cudaEventRecord(ctx->ev[2], st));
  {
#define ASM(XP, NEN, F) k_assemble<XP, NEN, F><<<(unsigned)g.ntiles, TILE_VOX, 0, st>>>(g, kz0, kz1, ctx->tile_faces.as<unsigned char>(), ctx->tile_ptr.as<int>(), sorted, \
        ctx->act_rec.as<ActRec>(), ctx->tri_rec.as<TriRec>(), ctx->IEN32.as<int>(), ctx->X.as<double>(), ctx->rho_n.as<double>(), rho_t, delta, ctx->pairbuf.as<double>(), ctx->pairxp.as<double>(), \
        ctx->dist.as<double>(), ctx->xp.as<double>()); LAUNCH_CHECK()
    if (minpath || atom) { ASM(false, 8, true); }
    else if (nen == 8) { if (want_xp) { ASM(true, 8, false); ASM(true, 8, true); } else { ASM(false, 8, false); ASM(false, 8, true); } }
    else { if (want_xp) { ASM(true, 4, false); ASM(true, 4, true); } else { ASM(false, 4, false); ASM(false, 4, true); } }
#undef ASM
  }
  CK(cudaEventRecord(ctx->ev[3], st));
  ctx->h_counters.resize(8 * 1025); u64 *hall = ctx->h_counters.data(); u64 hc[8];
  CK(cudaMemcpyAsync(hall, ctx->counters.p, sizeof(u64) * 8 * 1025, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  for (int q = 0; q < 8; q++) { hc[q] = hall[q]; for (int sl = 1; sl <= 1024; sl++) hc[q] += hall[8 * sl + q]; }
  ctx->rep.n_solid = (i64)hc[0]; ctx->rep.n_crossing = (i64)hc[1]; ctx->rep.n_active = nact; ctx->rep.n_pairs = npairs;
  ctx->rep.n_newton_iters = (i64)hc[2]; ctx->rep.n_not_converged = (i64)hc[3];
  CK(cudaEventElapsedTime(&ctx->rep.ms_bin, ctx->ev[0], ctx->ev[1]));
  CK(cudaEventElapsedTime(&ctx->rep.ms_project, ctx->ev[1], ctx->ev[2]));
  CK(cudaEventElapsedTime(&ctx->rep.ms_assemble, ctx->ev[2], ctx->ev[3]));
  return 0;
}
