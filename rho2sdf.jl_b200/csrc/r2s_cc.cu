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
// Components cross slab boundaries, but the only input of the labelling is the 1-bit interior mask: every rank packs the
// mask of its planes (1 bit per grid point), one all-gather replicates the whole mask (ngp / 8 bytes -- 17 MB at 519^3),
// every rank labels the full grid redundantly and flips only its own planes.  No label merging across ranks, bit-identical
// to the single-GPU result by construction.
struct SlabMap { int nranks; int k0[65]; };
__global__ void k_mask_pack(int nxy, int wpp, int kz0, int kz1, const double *__restrict__ sdf, double thr, unsigned *__restrict__ bits) {
  // one warp per 32 consecutive points of a plane
  i64 wid = (blockIdx.x * (i64)blockDim.x + threadIdx.x) >> 5; int lane = threadIdx.x & 31;
  i64 nw = (i64)wpp * (kz1 - kz0);
  if (wid >= nw) return;
  int pl = (int)(wid / wpp), w = (int)(wid % wpp), idx = w * 32 + lane;
  bool in = idx < nxy && sdf[(i64)(kz0 + pl) * nxy + idx] >= thr;
  unsigned m = __ballot_sync(0xffffffffu, in);
  if (lane == 0) bits[wid] = m;
}
__global__ void k_cc_init_bits(int nxy, int nx, int nz, int wpp, i64 stride, SlabMap sm, const unsigned *__restrict__ bits, int *__restrict__ L, int *__restrict__ sz) {
  i64 v = blockIdx.x * (i64)blockDim.x + threadIdx.x;
  const i64 n = (i64)nxy * nz;
  bool interior = false;
  if (v < n) {
    int k = (int)(v / nxy), idx = (int)(v % nxy), r = 0;
    while (r + 1 < sm.nranks && k >= sm.k0[r + 1]) r++;
    unsigned word = bits[(i64)r * stride + (i64)(k - sm.k0[r]) * wpp + (idx >> 5)];
    interior = (word >> (idx & 31)) & 1u;
  }
  int lab = run_start_label(v < n ? v : n, nx, interior);
  if (v >= n) return;
  L[v] = lab; sz[v] = 0;
}
static int remove_artifacts_slabs(r2s_ctx *ctx, double thr, double ratio, i64 *flipped) {
  const GridDev &g = ctx->g; cudaStream_t st = ctx->stream;
  const int nx = g.np[0], ny = g.np[1], nz = g.np[2], nxy = nx * ny, wpp = (nxy + 31) / 32;
  const i64 n = (i64)nxy * nz;
  if (n >= (1ll << 31)) FAIL("remove_sdf_artifacts: grid too large for 32-bit labels");
  if ((int)ctx->slab_k0.size() != ctx->nranks + 1) FAIL("remove_sdf_artifacts: slab table missing (call r2s_set_slab after r2s_comm_init)");
  SlabMap sm; sm.nranks = ctx->nranks; int maxpl = 0;
  if (ctx->nranks > 64) FAIL("remove_sdf_artifacts: more than 64 slabs");
  for (int r = 0; r <= ctx->nranks; r++) sm.k0[r] = ctx->slab_k0[r];
  for (int r = 0; r < ctx->nranks; r++) maxpl = std::max(maxpl, sm.k0[r + 1] - sm.k0[r]);
  const i64 stride = (i64)maxpl * wpp;
  CK(ctx->cc_bits_all.reserve(sizeof(unsigned) * (size_t)(stride * ctx->nranks)));
  CK(ctx->cc_label.reserve(sizeof(int) * (size_t)n));
  CK(ctx->cc_size.reserve(sizeof(int) * (size_t)n));
  CK(ctx->cc_scal.reserve(sizeof(u64) * 4));
  CK(cudaMemsetAsync(ctx->cc_scal.p, 0, sizeof(u64) * 4, st));
  unsigned *all = ctx->cc_bits_all.as<unsigned>(), *mine = all + stride * ctx->rank;
  int *L = ctx->cc_label.as<int>(), *sz = ctx->cc_size.as<int>(); u64 *sc = ctx->cc_scal.as<u64>();
  double *sdf = ctx->sdf.as<double>();
  const int kz0 = (int)ctx->k0, kz1 = (int)ctx->k1;
  CK(cudaMemsetAsync(mine, 0, sizeof(unsigned) * (size_t)stride, st));
  k_mask_pack<<<cdiv((i64)wpp * (kz1 - kz0) * 32, 256), 256, 0, st>>>(nxy, wpp, kz0, kz1, sdf, thr, mine); LAUNCH_CHECK();
  if (r2s_allgather_u32(ctx, mine, all, (size_t)stride)) return 1;       // in place: my block already sits at its slot
  int nb = cdiv(n, 256);
  k_cc_init_bits<<<nb, 256, 0, st>>>(nxy, nx, nz, wpp, stride, sm, all, L, sz); LAUNCH_CHECK();
  k_cc_merge<<<nb, 256, 0, st>>>(nx, ny, nz, L); LAUNCH_CHECK();
  k_cc_flatten<<<nb, 256, 0, st>>>(n, L, sz); LAUNCH_CHECK();
  k_cc_largest<<<nb, 256, 0, st>>>(n, L, sz, sc); LAUNCH_CHECK();
  u64 best = 0;
  CK(cudaMemcpyAsync(&best, sc, sizeof(u64), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *flipped = 0;
  if (best == 0) return 0;
  i64 largest = (i64)(best >> 32); int lroot = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
  double q = ratio * (double)largest; i64 ms = (i64)nearbyint(q); if (ms < 1) ms = 1;
  if (ms > 0x7fffffff) ms = 0x7fffffff;
  // flip my planes only: labels are global indices, so offset both arrays to the slab
  i64 v0 = (i64)kz0 * nxy, nloc = (i64)(kz1 - kz0) * nxy;
  k_cc_flip<<<cdiv(nloc, 256), 256, 0, st>>>(nloc, v0, L + v0, sz, lroot, (int)ms, sdf, sc + 1); LAUNCH_CHECK();
  if (r2s_allreduce(ctx, sc + 1, 1, 1)) return 1;
  u64 nf = 0;
  CK(cudaMemcpyAsync(&nf, sc + 1, sizeof(u64), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
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
  CK(cudaMemcpyAsync(&best, sc, sizeof(u64), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *flipped = 0;
  if (best == 0) return 0;                                          // no interior nodes (:149-152)
  i64 largest = (i64)(best >> 32); int lroot = (int)(0xffffffffu - (unsigned)(best & 0xffffffffu));
  // min_component_size = max(1, round(Int, ratio * largest)), round-half-to-even (:206)
  double q = ratio * (double)largest; i64 ms = (i64)nearbyint(q); if (ms < 1) ms = 1;
  if (ms > 0x7fffffff) ms = 0x7fffffff;
  k_cc_flip<<<nb, 256, 0, st>>>(n, v0, L, sz, lroot, (int)ms, sdf, sc + 1); LAUNCH_CHECK();
  u64 nf = 0;
  CK(cudaMemcpyAsync(&nf, sc + 1, sizeof(u64), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  *flipped = (i64)nf;
  return 0;
}
