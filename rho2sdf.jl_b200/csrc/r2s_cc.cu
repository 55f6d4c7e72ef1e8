// r2s_cc.cu -- remove_sdf_artifacts! (SignedDistances/SdfArtifactRemoval.jl:134-245) on the GPU
//
// 6-connected components of {sdf >= threshold} by a lock-free union-find (atomicCAS hooking of the larger root under
// the smaller one, path halving), canonical label = smallest linear index of the component.  Labels are not an output
// of the reference -- only the set of flipped points is -- so any canonical labelling is parity-safe; the largest
// component is chosen by (size, then smallest root), which is deterministic.
#include <algorithm>
#include "r2s_common.cuh"

__device__ __forceinline__ int uf_find(int *L, int x) {
  while (true) {
    int p = L[x];
    if (p == x) return x;
    int gp = L[p];
    if (gp != p) L[x] = gp;     // path halving (benign race: only ever replaces a parent by an ancestor)
    x = p;
  }
}
__device__ __forceinline__ void uf_union(int *L, int a, int b) {
  while (true) {
    a = uf_find(L, a); b = uf_find(L, b);
    if (a == b) return;
    if (a < b) { int t = a; a = b; b = t; }      // a > b: hook a under b
    int old = atomicCAS(&L[a], a, b);
    if (old == a) return;
  }
}
// Labels start as x-RUN STARTS: an interior voxel points at the first voxel of its contiguous interior run inside the warp's
// 32 consecutive voxels of the same grid row (ballot + clz), so that whole runs are one tree from the start and the merge
// kernel only has to link runs: across the warp boundary in x, and to the rows above (y) / planes above (z) once per pair of
// overlapping runs instead of once per voxel pair.
__device__ __forceinline__ int run_start_label(i64 v, int nx, bool interior) {
  const int lane = threadIdx.x & 31;
  const i64 row = v / nx;
  const unsigned same = __match_any_sync(0xffffffffu, row);
  const unsigned ones = __ballot_sync(0xffffffffu, interior) & same;
  const unsigned z = ~ones & ((1u << lane) - 1u);
  const int start = z ? 32 - __clz(z) : 0;
  return interior ? (int)(v - (lane - start)) : -1;
}
__global__ void k_cc_init(i64 n, i64 v0, int nx, const double *__restrict__ sdf, double thr, int *__restrict__ L, int *__restrict__ sz) {
  i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  bool interior = v < n && sdf[v0 + v] >= thr;
  int lab = run_start_label(v < n ? v : n, nx, interior);      // out-of-range lanes form their own "row"
  if (v >= n) return;
  L[v] = lab; sz[v] = 0;
}
__global__ void k_cc_merge(int nx, int ny, int nz, int *__restrict__ L) {
  const i64 n = (i64)nx * ny * nz, pl = (i64)nx * ny;
  const i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (v >= n || L[v] < 0) return;
  const int i = (int)(v % nx), j = (int)((v / nx) % ny), k = (int)(v / pl);
  const bool left_in = i > 0 && L[v - 1] >= 0;
  // x: runs were cut at warp boundaries
  if (left_in && (threadIdx.x & 31) == 0) uf_union(L, (int)v, (int)v - 1);
  // y / z: one link per pair of overlapping runs -- at the first overlapping x either this voxel or the neighbour starts its run
  if (j + 1 < ny && L[v + nx] >= 0 && (!left_in || !(L[v + nx - 1] >= 0))) uf_union(L, (int)v, (int)(v + nx));
  if (k + 1 < nz && L[v + pl] >= 0 && (!left_in || !(L[v + pl - 1] >= 0))) uf_union(L, (int)v, (int)(v + pl));
}
__global__ void k_cc_flatten(i64 n, int *__restrict__ L, int *__restrict__ sz) {
  i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (v >= n || L[v] < 0) return;
  // read-only walk to the root: a path-halving write here could land AFTER another thread stored its final root
  // label and replace it by a non-root ancestor
  int r = (int)v;
  while (true) { int p = L[r]; if (p == r) break; r = p; }
  L[v] = r;
  // one atomic per distinct root per warp (the largest component would otherwise serialise on a single address)
  unsigned act = __activemask();
  unsigned peers = __match_any_sync(act, r);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sz[r], __popc(peers));
}
__global__ void k_cc_largest(i64 n, const int *__restrict__ L, const int *__restrict__ sz, u64 *__restrict__ best) {
  i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  u64 key = 0;
  if (v < n && L[v] == (int)v) key = ((u64)(unsigned)sz[v] << 32) | (u64)(0xffffffffu - (unsigned)v);
  for (int o = 16; o > 0; o >>= 1) { u64 other = __shfl_down_sync(0xffffffffu, key, o); if (other > key) key = other; }
  if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}
__global__ void k_cc_flip(i64 n, i64 v0, const int *__restrict__ L, const int *__restrict__ sz, int lroot, int min_size, double *__restrict__ sdf, u64 *__restrict__ nflip) {
  i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  bool flip = false;
  if (v < n) { int r = L[v]; flip = (r >= 0 && r != lroot && sz[r] < min_size); }
  if (flip) sdf[v0 + v] = -fabs(sdf[v0 + v]);
  unsigned m = __ballot_sync(0xffffffffu, flip);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nflip, (u64)__popc(m));
}

// ---- multi-rank (z-slab) variant ----------------------------------------------------------------------------------
// Every rank labels ITS planes (labels are global grid-point indices, arrays are globally indexed), then the components
// that touch a slab boundary are stitched on a small replicated graph:
//   1. the labels (= local roots) and local sizes of each rank's bottom and top plane are all-gathered (4 planes of ints/rank);
//   2. every rank enters the foreign roots into its own parent array and unions the label pairs facing each other across
//      every interface -- the same deterministic min-index forest on every rank;
//   3. the size of a stitched component is the sum of the local sizes of its member roots (each counted once);
//   4. largest component = max over the stitched components (known everywhere) and each rank's interior components
//      (one 64-bit all-reduce of the (size, ~root) key) -- same tie-break as the single-GPU path;
//   5. every rank flips its own voxels.
// The flipped set is identical to the single-GPU result (canonical root = smallest index of the component).
#define CC_MAXR 64
__device__ __forceinline__ int find_ro(const int *L, int x) { while (true) { int p = L[x]; if (p == x) return x; x = p; } }
__global__ void k_ccg_init(i64 nloc, i64 v0, int nx, const double *__restrict__ sdf, double thr, int *__restrict__ L, int *__restrict__ sz, int *__restrict__ seen) {
  i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  bool interior = v < nloc && sdf[v0 + v] >= thr;
  int lab = run_start_label(v < nloc ? v0 + v : v0 + nloc, nx, interior);      // v0 is a multiple of nx: rows are intact
  if (v >= nloc) return;
  L[v0 + v] = lab; sz[v0 + v] = 0; seen[v0 + v] = 0;
}
__global__ void k_ccg_merge(int nx, int ny, int kz0, int kz1, int *__restrict__ L) {
  const i64 pl = (i64)nx * ny, nloc = pl * (kz1 - kz0);
  const i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nloc) return;
  const i64 v = pl * kz0 + t;
  if (L[v] < 0) return;
  const int i = (int)(v % nx), j = (int)((v / nx) % ny), k = (int)(v / pl);
  const bool left_in = i > 0 && L[v - 1] >= 0;
  if (left_in && (threadIdx.x & 31) == 0) uf_union(L, (int)v, (int)v - 1);
  if (j + 1 < ny && L[v + nx] >= 0 && (!left_in || !(L[v + nx - 1] >= 0))) uf_union(L, (int)v, (int)(v + nx));
  if (k + 1 < kz1 && L[v + pl] >= 0 && (!left_in || !(L[v + pl - 1] >= 0))) uf_union(L, (int)v, (int)(v + pl));
}
__global__ void k_ccg_flatten(i64 nloc, i64 v0, int *__restrict__ L, int *__restrict__ sz) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= nloc || L[v0 + t] < 0) return;
  int r = find_ro(L, (int)(v0 + t));
  L[v0 + t] = r;
  unsigned act = __activemask();
  unsigned peers = __match_any_sync(act, r);
  if ((threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&sz[r], __popc(peers));
}
// boundary planes of this rank: [bottom labels | top labels | bottom sizes | top sizes], nxy ints each
__global__ void k_ccg_planes(int nxy, i64 vbot, i64 vtop, const int *__restrict__ L, const int *__restrict__ sz, int *__restrict__ out) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= nxy) return;
  int a = L[vbot + p], b = L[vtop + p];
  out[p] = a; out[nxy + p] = b; out[2 * nxy + p] = a >= 0 ? sz[a] : 0; out[3 * nxy + p] = b >= 0 ? sz[b] : 0;
}
// enter every boundary root (own and foreign) into the replicated graph
__global__ void k_ccg_scatter(int nxy, int nranks, i64 own_lo, i64 own_hi, const int *__restrict__ all, int *__restrict__ L, int *__restrict__ sz, int *__restrict__ gsz,
                              int *__restrict__ seen) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= (i64)nranks * 2 * nxy) return;
  int r = (int)(t / (2 * nxy)), q = (int)(t % (2 * nxy));
  const int *blk = all + (i64)r * 4 * nxy;
  int a = blk[q];
  if (a < 0) return;
  if (a < own_lo || a >= own_hi) { L[a] = a; sz[a] = blk[2 * nxy + q]; }
  gsz[a] = 0; seen[a] = 0;
}
__global__ void k_ccg_union(int nxy, int nranks, const int *__restrict__ all, int *__restrict__ L) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  int a = -1, b = -1;
  if (t < (i64)(nranks - 1) * nxy) {
    int r = (int)(t / nxy), p = (int)(t % nxy);
    a = all[(i64)r * 4 * nxy + nxy + p]; b = all[(i64)(r + 1) * 4 * nxy + p];      // top of r, bottom of r + 1
  }
  // the big components face each other along thousands of voxels: one union per distinct (a, b) pair per warp
  const bool on = a >= 0 && b >= 0;
  const u64 key = on ? (((u64)(unsigned)a << 32) | (unsigned)b) : ~0ull;
  const unsigned peers = __match_any_sync(0xffffffffu, key);
  if (on && (threadIdx.x & 31) == __ffs(peers) - 1) uf_union(L, a, b);
}
__global__ void k_ccg_sizes(int nxy, int nranks, const int *__restrict__ all, const int *__restrict__ L, const int *__restrict__ sz, int *__restrict__ gsz, int *__restrict__ seen) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  if (t >= (i64)nranks * 2 * nxy) return;
  int r = (int)(t / (2 * nxy)), q = (int)(t % (2 * nxy));
  int a = all[(i64)r * 4 * nxy + q];
  const unsigned peers = __match_any_sync(__activemask(), a);
  if (a < 0 || (threadIdx.x & 31) != __ffs(peers) - 1) return;      // one lane per distinct root per warp
  if (atomicExch(&seen[a], 1) == 0) atomicAdd(&gsz[find_ro(L, a)], sz[a]);
}
__global__ void k_ccg_largest(int nxy, int nranks, i64 nloc, i64 v0, const int *__restrict__ all, const int *__restrict__ L, const int *__restrict__ sz, const int *__restrict__ gsz,
                              const int *__restrict__ seen, u64 *__restrict__ best) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  u64 key = 0;
  const i64 nb = (i64)nranks * 2 * nxy;
  if (t < nb) {                      // stitched components (identical on every rank)
    int r = (int)(t / (2 * nxy)), q = (int)(t % (2 * nxy));
    int a = all[(i64)r * 4 * nxy + q];
    if (a >= 0) { int g = find_ro(L, a); key = ((u64)(unsigned)gsz[g] << 32) | (u64)(0xffffffffu - (unsigned)g); }
  } else if (t < nb + nloc) {        // my components that touch no slab boundary
    i64 v = v0 + (t - nb);
    if (L[v] == (int)v && seen[v] == 0) key = ((u64)(unsigned)sz[v] << 32) | (u64)(0xffffffffu - (unsigned)v);
  }
  for (int o = 16; o > 0; o >>= 1) { u64 other = __shfl_down_sync(0xffffffffu, key, o); if (other > key) key = other; }
  if ((threadIdx.x & 31) == 0 && key) atomicMax(best, key);
}
__global__ void k_ccg_flip(i64 nloc, i64 v0, const int *__restrict__ L, const int *__restrict__ sz, const int *__restrict__ gsz, const int *__restrict__ seen, int lroot, int min_size,
                           double *__restrict__ sdf, u64 *__restrict__ nflip) {
  i64 t = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  bool flip = false;
  if (t < nloc) {
    int a = L[v0 + t];
    if (a >= 0) {
      int g = a, size = sz[a];
      if (seen[a]) { g = find_ro(L, a); size = gsz[g]; }
      flip = g != lroot && size < min_size;
    }
  }
  if (flip) sdf[v0 + t] = -fabs(sdf[v0 + t]);
  unsigned m = __ballot_sync(0xffffffffu, flip);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(nflip, (u64)__popc(m));
}
static int remove_artifacts_slabs(r2s_ctx *ctx, double thr, double ratio, i64 *flipped) {
  const GridDev &g = ctx->g; cudaStream_t st = ctx->stream;
  const int nx = g.np[0], ny = g.np[1], nz = g.np[2], nxy = nx * ny, R = ctx->nranks;
  const i64 n = (i64)nxy * nz;
  if (n >= (1ll << 31)) FAIL("remove_sdf_artifacts: grid too large for 32-bit labels");
  if (R > CC_MAXR) FAIL("remove_sdf_artifacts: more than 64 slabs");
  const int kz0 = (int)ctx->k0, kz1 = (int)ctx->k1;
  const i64 v0 = (i64)kz0 * nxy, nloc = (i64)(kz1 - kz0) * nxy;
  CK(ctx->cc_label.reserve(sizeof(int) * (size_t)n)); CK(ctx->cc_size.reserve(sizeof(int) * (size_t)n));
  CK(ctx->cc_gsz.reserve(sizeof(int) * (size_t)n)); CK(ctx->cc_seen.reserve(sizeof(int) * (size_t)n));
  CK(ctx->cc_bits_all.reserve(sizeof(int) * (size_t)4 * nxy * R));
  CK(ctx->cc_scal.reserve(sizeof(u64) * 4));
  CK(cudaMemsetAsync(ctx->cc_scal.p, 0, sizeof(u64) * 4, st));
  int *L = ctx->cc_label.as<int>(), *sz = ctx->cc_size.as<int>(), *gsz = ctx->cc_gsz.as<int>(), *seen = ctx->cc_seen.as<int>(), *all = ctx->cc_bits_all.as<int>();
  u64 *sc = ctx->cc_scal.as<u64>(); double *sdf = ctx->sdf.as<double>();
  const int nbl = cdiv(nloc, 256);
  k_ccg_init<<<nbl, 256, 0, st>>>(nloc, v0, nx, sdf, thr, L, sz, seen); LAUNCH_CHECK();
  k_ccg_merge<<<nbl, 256, 0, st>>>(nx, ny, kz0, kz1, L); LAUNCH_CHECK();
  k_ccg_flatten<<<nbl, 256, 0, st>>>(nloc, v0, L, sz); LAUNCH_CHECK();
  int *mine = all + (i64)ctx->rank * 4 * nxy;
  k_ccg_planes<<<cdiv(nxy, 256), 256, 0, st>>>(nxy, v0, v0 + nloc - nxy, L, sz, mine); LAUNCH_CHECK();
  if (r2s_allgather_u32(ctx, (const unsigned *)mine, (unsigned *)all, (size_t)4 * nxy)) return 1;      // in place
  const i64 nb = (i64)R * 2 * nxy;
  k_ccg_scatter<<<cdiv(nb, 256), 256, 0, st>>>(nxy, R, v0, v0 + nloc, all, L, sz, gsz, seen); LAUNCH_CHECK();
  k_ccg_union<<<cdiv((i64)(R - 1) * nxy, 256), 256, 0, st>>>(nxy, R, all, L); LAUNCH_CHECK();
  k_ccg_sizes<<<cdiv(nb, 256), 256, 0, st>>>(nxy, R, all, L, sz, gsz, seen); LAUNCH_CHECK();
  k_ccg_largest<<<cdiv(nb + nloc, 256), 256, 0, st>>>(nxy, R, nloc, v0, all, L, sz, gsz, seen, sc); LAUNCH_CHECK();
  if (r2s_allreduce(ctx, sc, 1, 4)) return 1;
  u64 best = 0;
  if (r2s_readback(ctx, &best, sc, sizeof(u64))) return 1;
  *flipped = 0;
  if (best == 0) return 0;
  i64 largest = (i64)(best >> 32); int lroot = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
  double q = ratio * (double)largest; i64 ms = (i64)nearbyint(q); if (ms < 1) ms = 1;
  if (ms > 0x7fffffff) ms = 0x7fffffff;
  k_ccg_flip<<<nbl, 256, 0, st>>>(nloc, v0, L, sz, gsz, seen, lroot, (int)ms, sdf, sc + 1); LAUNCH_CHECK();
  if (r2s_allreduce(ctx, sc + 1, 1, 1)) return 1;
  u64 nf = 0;
  if (r2s_readback(ctx, &nf, sc + 1, sizeof(u64))) return 1;
  *flipped = (i64)nf;
  return 0;
}

int r2s_dev_remove_artifacts(r2s_ctx *ctx, double thr, double ratio, i64 *flipped) {
  if (ctx->nranks > 1) return remove_artifacts_slabs(ctx, thr, ratio, flipped);
  const GridDev &g = ctx->g; cudaStream_t st = ctx->stream;
  int nx = g.np[0], ny = g.np[1], nz = (int)(ctx->k1 - ctx->k0);
  i64 n = (i64)nx * ny * nz, v0 = (i64)ctx->k0 * nx * ny;
  if (n >= (1ll << 31)) FAIL("remove_sdf_artifacts: slab too large for 32-bit labels");
  CK(ctx->cc_label.reserve(sizeof(int) * (size_t)n));
  CK(ctx->cc_size.reserve(sizeof(int) * (size_t)n));
  CK(ctx->cc_scal.reserve(sizeof(u64) * 4));
  CK(cudaMemsetAsync(ctx->cc_scal.p, 0, sizeof(u64) * 4, st));
  int *L = ctx->cc_label.as<int>(), *sz = ctx->cc_size.as<int>(); u64 *sc = ctx->cc_scal.as<u64>();
  double *sdf = ctx->sdf.as<double>();
  int nb = cdiv(n, 256);
  k_cc_init<<<nb, 256, 0, st>>>(n, v0, nx, sdf, thr, L, sz); LAUNCH_CHECK();
  k_cc_merge<<<nb, 256, 0, st>>>(nx, ny, nz, L); LAUNCH_CHECK();
  k_cc_flatten<<<nb, 256, 0, st>>>(n, L, sz); LAUNCH_CHECK();
  k_cc_largest<<<nb, 256, 0, st>>>(n, L, sz, sc); LAUNCH_CHECK();
  u64 best = 0;
  if (r2s_readback(ctx, &best, sc, sizeof(u64))) return 1;
  *flipped = 0;
  if (best == 0) return 0;                                          // no interior nodes (:149-152)
  i64 largest = (i64)(best >> 32); int lroot = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
  // min_component_size = max(1, round(Int, ratio * largest)), round-half-to-even (:206)
  double q = ratio * (double)largest; i64 ms = (i64)nearbyint(q); if (ms < 1) ms = 1;
  if (ms > 0x7fffffff) ms = 0x7fffffff;
  k_cc_flip<<<nb, 256, 0, st>>>(n, v0, L, sz, lroot, (int)ms, sdf, sc + 1); LAUNCH_CHECK();
  u64 nf = 0;
  if (r2s_readback(ctx, &nf, sc + 1, sizeof(u64))) return 1;
  *flipped = (i64)nf;
  return 0;
}
