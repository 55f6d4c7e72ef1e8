// r2s_comm.cu -- z-slab exchange layer: NCCL over NVLink 5 / NVSwitch, one process per GPU.
//
// The reference has no distributed path.  Here the coarse SDF grid is cut into contiguous z-slabs, one per rank; distances
// and signs need no communication (elements straddling a slab boundary are processed by both ranks); the exchanges are
//   * halo planes of the Float32 smoothing fields (ncclSend/ncclRecv with the z-neighbours, grouped),
//   * scalar all-reduces (CG dot products, max/min, volume partial sums),
//   * one all-gather of the 1-bit interior mask for the artifact removal.
// NCCL is bound at run time with dlopen so that libr2s.so loads (and the single-GPU path works) where NCCL is absent; in
// a process that already has NCCL loaded (torch) the same library instance is reused.  The host program only has to carry
// the 128-byte ncclUniqueId from rank 0 to the other ranks (torch.distributed broadcast / MPI / a file).
#include <dlfcn.h>
#include <stdlib.h>
#include <string.h>
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <vector>
#include "r2s_common.cuh"

typedef struct { char internal[128]; } nccl_uid;
typedef void *nccl_comm;
enum { NC_INT8 = 0, NC_UINT8 = 1, NC_INT32 = 2, NC_UINT32 = 3, NC_INT64 = 4, NC_UINT64 = 5, NC_F16 = 6, NC_F32 = 7, NC_F64 = 8 };
enum { NC_SUM = 0, NC_PROD = 1, NC_MAX = 2, NC_MIN = 3 };

struct NcclApi {
  void *lib = nullptr;
  int (*GetUniqueId)(nccl_uid *);
  int (*CommInitRank)(nccl_comm *, int, nccl_uid, int);
  int (*CommDestroy)(nccl_comm);
  int (*AllReduce)(const void *, void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t);
  int (*Send)(const void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, nccl_comm, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char *(*GetErrorString)(int);
};
static NcclApi g_nccl;
// ---- in-process slab group -------------------------------------------------------------------------------------------------
// r2s_multi_* (r2s_multi.cu) runs one context per slab inside ONE process, each driven by its own host thread.  The ranks then
// see each other's device memory directly (peer access between devices, plain pointers on one device), so nothing here needs NCCL
// or CUDA IPC: scalar all-reduces and the CG halo planes use the same mailbox kernels as the multi-process path, bulk exchanges
// (halo planes of the smoothing fields, the all-gather of the artifact removal) are cudaMemcpyPeerAsync pulls ordered by events,
// with two host barriers per exchange ("source ready" / "everybody has read").  Several slabs may share one GPU (tests on a 1-GPU box).
struct LocalGroup {
  int n = 0; r2s_ctx *ctx[64];
  std::mutex mu; std::condition_variable cv; int arrived = 0; unsigned long gen = 0; bool failed = false;
  const void *src[64]; size_t cnt[64]; cudaEvent_t ev_ready[64], ev_done[64];
  bool spin_ok = false;      // every slab on its own device: the device-side mailbox (spin waits) is safe.  Slabs that SHARE a device must not wait for
                             // each other inside kernels (an allocation or a full hardware queue on the shared device can hold back the kernel that is
                             // waited for), so they use the event-ordered exchanges below for everything.
  void *stage[64];           // per slab: 64 x 4 words of staging for the event-ordered all-reduce
  // returns false when some rank has failed (the caller must bail out instead of waiting for it)
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (failed) return false;
    const unsigned long my = gen;
    if (++arrived == n) { arrived = 0; gen++; cv.notify_all(); return true; }
    cv.wait(lk, [&] { return gen != my || failed; });
    return !failed;
  }
};
void r2s_local_group_abort(LocalGroup *g) { if (!g) return; std::lock_guard<std::mutex> lk(g->mu); g->failed = true; g->cv.notify_all(); }
#define LBAR(lg) do { if (!(lg)->barrier()) FAIL("in-process slab group: another slab failed"); } while (0)
static int p2p_setup(r2s_ctx *ctx);
static void p2p_teardown(r2s_ctx *ctx);
static const char *nccl_load() {
  if (g_nccl.lib) return nullptr;
  const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
  void *h = nullptr;
  for (int i = 0; names[i] && !h; i++) h = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
  if (!h) return "libnccl.so.2 could not be loaded (multi-GPU slabs need NCCL)";
#define SYM(field, name) *(void **)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) return "NCCL symbol missing: " name;
  SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(AllReduce, "ncclAllReduce")
  SYM(AllGather, "ncclAllGather") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv") SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  g_nccl.lib = h;
  return nullptr;
}
#define NCK(call)                                                                              \
  do {                                                                                         \
    int _r = (call);                                                                           \
    if (_r != 0) {                                                                             \
      char _b[512];                                                                            \
      snprintf(_b, sizeof(_b), "%s:%d: %s -> NCCL: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(_r)); \
      ctx->err = _b;                                                                           \
      return 1;                                                                                \
    }                                                                                          \
  } while (0)

extern "C" int r2s_comm_unique_id(void *id128) {
  if (!id128 || nccl_load()) return 1;
  nccl_uid id;
  if (g_nccl.GetUniqueId(&id) != 0) return 2;
  memcpy(id128, &id, sizeof(id));
  return 0;
}
extern "C" int r2s_comm_init(r2s_ctx *ctx, int rank, int nranks, const void *id128) {
  if (!ctx) return 1;
  if (nranks < 1 || rank < 0 || rank >= nranks || !id128) FAIL("r2s_comm_init: invalid rank / size");
  if (const char *e = nccl_load()) FAIL(e);
  CK(cudaSetDevice(ctx->device));
  nccl_uid id; memcpy(&id, id128, sizeof(id));
  nccl_comm c = nullptr;
  NCK(g_nccl.CommInitRank(&c, nranks, id, rank));
  ctx->comm = c; ctx->rank = rank; ctx->nranks = nranks;
  return p2p_setup(ctx);
}
extern "C" int r2s_comm_destroy(r2s_ctx *ctx) {
  if (!ctx) return 1;
  if (ctx->lg) { cudaSetDevice(ctx->device); cudaStreamSynchronize(ctx->stream); if (ctx->p2p_peer_box_dev) cudaFree(ctx->p2p_peer_box_dev); if (ctx->p2p_box) cudaFree(ctx->p2p_box);
                 ctx->p2p_box = nullptr; ctx->p2p_peer_box_dev = nullptr; ctx->p2p = false; ctx->p2p_c_local = nullptr; ctx->p2p_c_peer[0] = ctx->p2p_c_peer[1] = nullptr; ctx->lg = nullptr; }
  if (ctx->comm && g_nccl.lib) { cudaStreamSynchronize(ctx->stream); p2p_teardown(ctx); g_nccl.CommDestroy((nccl_comm)ctx->comm); }
  ctx->comm = nullptr; ctx->rank = 0; ctx->nranks = 1;
  return 0;
}

// fuse the collectives issued between the two calls into one NCCL launch (nested groups are allowed)
int r2s_group_start(r2s_ctx *ctx) { if (ctx->nranks > 1 && !ctx->lg) NCK(g_nccl.GroupStart()); return 0; }
int r2s_group_end(r2s_ctx *ctx) { if (ctx->nranks > 1 && !ctx->lg) NCK(g_nccl.GroupEnd()); return 0; }
// ---- collectives used by the pipeline; all are no-ops for a single rank ----------------------------------------------
static int p2p_allreduce(r2s_ctx *ctx, void *buf, size_t count, int kind);
static int local_allreduce(r2s_ctx *ctx, void *buf, size_t count, int kind);
int r2s_allreduce(r2s_ctx *ctx, void *buf, size_t count, int kind /*0 f64 sum, 1 u64 sum, 2 u32 max, 3 u32 min, 4 u64 max*/) {
  if (ctx->nranks <= 1) return 0;
  if (ctx->lg && count > 4) FAIL("in-process slab group: all-reduce of more than 4 words");
  if (ctx->lg && !ctx->p2p) return local_allreduce(ctx, buf, count, kind);
  if (ctx->p2p && count <= 4) return p2p_allreduce(ctx, buf, count, kind);
  int dt = kind == 0 ? NC_F64 : (kind == 1 || kind == 4 ? NC_UINT64 : NC_UINT32);
  int op = (kind == 0 || kind == 1) ? NC_SUM : (kind == 3 ? NC_MIN : NC_MAX);
  NCK(g_nccl.AllReduce(buf, buf, count, dt, op, (nccl_comm)ctx->comm, ctx->stream));
  ctx->collectives++;
  return 0;
}
// local (in-process) bulk exchanges: pull from the peers' buffers between two host barriers
static int local_pull_begin(r2s_ctx *ctx, const void *my_src, size_t my_bytes) {
  LocalGroup *lg = ctx->lg;
  lg->src[ctx->rank] = my_src; lg->cnt[ctx->rank] = my_bytes;
  CK(cudaEventRecord(lg->ev_ready[ctx->rank], ctx->stream));      // everything this rank has enqueued so far (its source data) precedes the peers' pulls
  LBAR(lg);
  return 0;
}
static int local_pull(r2s_ctx *ctx, int peer, void *dst, const void *src, size_t bytes) {
  LocalGroup *lg = ctx->lg;
  CK(cudaStreamWaitEvent(ctx->stream, lg->ev_ready[peer], 0));
  CK(cudaMemcpyPeerAsync(dst, ctx->device, src, lg->ctx[peer]->device, bytes, ctx->stream));
  return 0;
}
static int local_pull_end(r2s_ctx *ctx, const int *readers, int nreaders) {
  LocalGroup *lg = ctx->lg;
  CK(cudaEventRecord(lg->ev_done[ctx->rank], ctx->stream));       // my pulls are enqueued behind this point ...
  LBAR(lg);
  for (int i = 0; i < nreaders; i++) CK(cudaStreamWaitEvent(ctx->stream, lg->ev_done[readers[i]], 0));      // ... and my later writes wait for the ranks that read from me
  ctx->collectives++;
  return 0;
}
static int local_allgather(r2s_ctx *ctx, const void *send, void *recv, size_t bytes) {
  LocalGroup *lg = ctx->lg;
  if (local_pull_begin(ctx, send, bytes)) return 1;
  int readers[64], nr = 0;
  for (int q = 0; q < lg->n; q++) {
    char *dst = (char *)recv + (size_t)q * bytes;
    if (q == ctx->rank) { if (dst != (const char *)send) CK(cudaMemcpyAsync(dst, send, bytes, cudaMemcpyDeviceToDevice, ctx->stream)); }
    else { if (local_pull(ctx, q, dst, lg->src[q], bytes)) return 1; readers[nr++] = q; }
  }
  return local_pull_end(ctx, readers, nr);
}
// combine R contributions of n words in rank order (deterministic, identical on every slab)
__global__ void k_local_combine(const unsigned long long *__restrict__ stage, int R, int n, int kind, unsigned long long *__restrict__ out) {
  const int i = threadIdx.x;
  if (i >= n) return;
  unsigned long long acc = stage[i];
  for (int q = 1; q < R; q++) {
    const unsigned long long v = stage[q * 4 + i];
    if (kind == 0) acc = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)acc) + __longlong_as_double((long long)v));
    else if (kind == 1) acc += v;
    else if (kind == 3) acc = v < acc ? v : acc;
    else acc = v > acc ? v : acc;
  }
  out[i] = acc;
}
__global__ void k_local_combine32(const unsigned *__restrict__ stage, int R, int n, int kind, unsigned *__restrict__ out) {
  const int i = threadIdx.x;
  if (i >= n) return;
  unsigned acc = stage[i];
  for (int q = 1; q < R; q++) { const unsigned v = stage[q * 8 + i]; acc = kind == 3 ? (v < acc ? v : acc) : (v > acc ? v : acc); }
  out[i] = acc;
}
// event-ordered all-reduce of <= 4 words for slabs that share a device (no waiting inside kernels)
static int local_allreduce(r2s_ctx *ctx, void *buf, size_t count, int kind) {
  LocalGroup *lg = ctx->lg;
  const bool w32 = kind == 2 || kind == 3; const size_t es = w32 ? 4 : 8, slot = 32;      // 32 bytes per contribution
  if (local_pull_begin(ctx, buf, count * es)) return 1;
  char *stage = (char *)lg->stage[ctx->rank];
  int readers[64], nr = 0;
  for (int q = 0; q < lg->n; q++) {
    if (q == ctx->rank) CK(cudaMemcpyAsync(stage + (size_t)q * slot, buf, count * es, cudaMemcpyDeviceToDevice, ctx->stream));
    else { if (local_pull(ctx, q, stage + (size_t)q * slot, lg->src[q], count * es)) return 1; readers[nr++] = q; }
  }
  if (local_pull_end(ctx, readers, nr)) return 1;      // after this, nobody still reads my buf: it may be overwritten
  if (w32) k_local_combine32<<<1, 32, 0, ctx->stream>>>((const unsigned *)stage, lg->n, (int)count, kind, (unsigned *)buf);
  else k_local_combine<<<1, 32, 0, ctx->stream>>>((const unsigned long long *)stage, lg->n, (int)count, kind, (unsigned long long *)buf);
  CK(cudaGetLastError());
  return 0;
}
int r2s_allgather_u32(r2s_ctx *ctx, const unsigned *send, unsigned *recv, size_t count) {
  if (ctx->nranks <= 1) { CK(cudaMemcpyAsync(recv, send, count * sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream)); return 0; }
  if (ctx->lg) return local_allgather(ctx, send, recv, count * sizeof(unsigned));
  NCK(g_nccl.AllGather(send, recv, count, NC_UINT32, (nccl_comm)ctx->comm, ctx->stream));
  ctx->collectives++;
  return 0;
}
// Exchange halo planes of a globally indexed field a[plane * plane_elems ...] (4-byte elements): this rank owns planes
// [k0, k1) of nz (coarse or fine planes); after the call planes [k0 - below, k0) and [k1, k1 + above) hold the neighbours' values (clipped to [0, nz)).
int r2s_halo_exchange_f32(r2s_ctx *ctx, float *a, i64 plane_elems, int k0, int k1, int nz, int below, int above) {
  if (ctx->nranks <= 1) return 0;
  const int r = ctx->rank;
  if (ctx->lg) {      // in-process: pull the neighbours' boundary planes of the same (globally indexed) field
    LocalGroup *lg = ctx->lg;
    if (local_pull_begin(ctx, a, 0)) return 1;
    int readers[2], nrd = 0;
    if (r + 1 < ctx->nranks) {
      const int nr = above < nz - k1 ? above : nz - k1;      // planes [k1, k1 + nr) are the upper neighbour's first planes
      if (nr > 0 && local_pull(ctx, r + 1, a + (i64)k1 * plane_elems, (const float *)lg->src[r + 1] + (i64)k1 * plane_elems, sizeof(float) * (size_t)nr * plane_elems)) return 1;
      readers[nrd++] = r + 1;
    }
    if (r > 0) {
      const int nr = below < k0 ? below : k0;
      if (nr > 0 && local_pull(ctx, r - 1, a + (i64)(k0 - nr) * plane_elems, (const float *)lg->src[r - 1] + (i64)(k0 - nr) * plane_elems, sizeof(float) * (size_t)nr * plane_elems)) return 1;
      readers[nrd++] = r - 1;
    }
    return local_pull_end(ctx, readers, nrd);
  }
  NCK(g_nccl.GroupStart());
  if (r + 1 < ctx->nranks) {
    // the upper neighbour needs my top `below` planes; I need its bottom `above` planes
    int ns = below < k1 - k0 ? below : k1 - k0, nr = above < nz - k1 ? above : nz - k1;
    if (ns > 0) NCK(g_nccl.Send(a + (i64)(k1 - ns) * plane_elems, (size_t)ns * plane_elems, NC_F32, r + 1, (nccl_comm)ctx->comm, ctx->stream));
    if (nr > 0) NCK(g_nccl.Recv(a + (i64)k1 * plane_elems, (size_t)nr * plane_elems, NC_F32, r + 1, (nccl_comm)ctx->comm, ctx->stream));
  }
  if (r > 0) {
    int ns = above < k1 - k0 ? above : k1 - k0, nr = below < k0 ? below : k0;
    if (ns > 0) NCK(g_nccl.Send(a + (i64)k0 * plane_elems, (size_t)ns * plane_elems, NC_F32, r - 1, (nccl_comm)ctx->comm, ctx->stream));
    if (nr > 0) NCK(g_nccl.Recv(a + (i64)(k0 - nr) * plane_elems, (size_t)nr * plane_elems, NC_F32, r - 1, (nccl_comm)ctx->comm, ctx->stream));
  }
  NCK(g_nccl.GroupEnd());
  ctx->collectives++;
  return 0;
}

// ------------------------------------------------------------------------------------------------ peer-memory fast path
// NCCL costs 20-40 us per call even for 8 bytes; the CG needs two scalar all-reduces and one halo exchange per iteration and
// the threshold search one all-reduce per bisection, so at 8 GPUs those latencies were ~25 % of the step.  Here every rank
// owns a small MAILBOX in device memory that all peers map through CUDA IPC:
//   slots[parity][rank][8]  (8-byte words: 4 values, 1 sequence number)   and   halo flags[parity][2]
// All-reduce #s: one 64-thread kernel -- thread p stores this rank's values into slot[s&1][rank] of peer p's mailbox, fences
// (system scope) and stores s; then thread p spins on the local slot[s&1][p] until it carries s; thread 0 combines the R
// contributions IN RANK ORDER (deterministic, identical on every rank).  Two parities make it safe for a fast rank to start
// call s+1 while a slow one still combines call s (call s+2 cannot start before everybody finished call s).  The CG halo
// planes of c are written straight into the neighbours' c arrays (also IPC-mapped) BY THE MAT-VEC KERNEL ITSELF, whose last CTA also
// all-reduces the dot product through the mailbox and raises the halo flags; the update kernel waits for both neighbours' flags in
// its prologue (r2s_p2p.cuh, r2s_rbf.cu: fused compute + exchange, no separate communication launch inside a CG iteration).  Spins are bounded: a timeout raises an error
// flag instead of hanging the GPU.  R2S_P2P=0 falls back to NCCL for everything.
#include "r2s_p2p.cuh"
__global__ void k_p2p_allreduce(P2PBox *mine, P2PBox *const *peers, int rank, int R, unsigned seq, unsigned long long *vals, int n, int kind, const float *skip) {
  if (skip && *skip != 0.0f) return;      // every rank sees the same flag (it is computed from all-reduced values): all skip or none
  const int p = threadIdx.x, par = seq & 1;
  if (p < R) {
    volatile unsigned long long *dst = peers[p]->slot[par][rank];
    for (int i = 0; i < n; i++) dst[i] = vals[i];
    __threadfence_system();
    dst[P2P_SLOT_WORDS - 1] = (unsigned long long)seq;
  }
  if (p < R) {
    volatile unsigned long long *src = mine->slot[par][p];
    unsigned spin = 0;
    const unsigned lim = mine->error ? 1024u : P2P_MAX_SPIN;      // after a first time-out the later waits give up quickly
    while (src[P2P_SLOT_WORDS - 1] != (unsigned long long)seq) { if (++spin > lim) { mine->error = 1; break; } }
  }
  __threadfence_system();
  __syncthreads();
  if (p == 0) {
    for (int i = 0; i < n; i++) {
      volatile unsigned long long *s0 = mine->slot[par][0];
      unsigned long long acc = s0[i];
      for (int q = 1; q < R; q++) {
        unsigned long long v = ((volatile unsigned long long *)mine->slot[par][q])[i];
        if (kind == 0) acc = (unsigned long long)__double_as_longlong(__longlong_as_double((long long)acc) + __longlong_as_double((long long)v));
        else if (kind == 1) acc += v;
        else if (kind == 3) acc = v < acc ? v : acc;
        else acc = v > acc ? v : acc;                      // kinds 2, 4: max
      }
      vals[i] = acc;
    }
  }
}
// 32-bit kinds (2, 3) travel as zero-extended 64-bit words
__global__ void k_p2p_widen(const unsigned *in, unsigned long long *out, int n) { if ((int)threadIdx.x < n) out[threadIdx.x] = in[threadIdx.x]; }
__global__ void k_p2p_narrow(const unsigned long long *in, unsigned *out, int n) { if ((int)threadIdx.x < n) out[threadIdx.x] = (unsigned)in[threadIdx.x]; }

static int p2p_allreduce(r2s_ctx *ctx, void *buf, size_t count, int kind) {
  P2PBox *mine = (P2PBox *)ctx->p2p_box;
  unsigned long long *stage = (unsigned long long *)((char *)ctx->p2p_box + sizeof(P2PBox));      // 4 words behind the mailbox
  const unsigned seq = ++ctx->p2p_seq;
  if (kind == 2 || kind == 3) {
    k_p2p_widen<<<1, 32, 0, ctx->stream>>>((const unsigned *)buf, stage, (int)count);
    k_p2p_allreduce<<<1, 64, 0, ctx->stream>>>(mine, (P2PBox *const *)ctx->p2p_peer_box_dev, ctx->rank, ctx->nranks, seq, stage, (int)count, kind, ctx->skip_flag);
    k_p2p_narrow<<<1, 32, 0, ctx->stream>>>(stage, (unsigned *)buf, (int)count);
  } else {
    k_p2p_allreduce<<<1, 64, 0, ctx->stream>>>(mine, (P2PBox *const *)ctx->p2p_peer_box_dev, ctx->rank, ctx->nranks, seq, (unsigned long long *)buf, (int)count, kind, ctx->skip_flag);
  }
  CK(cudaGetLastError());
  ctx->p2p_ops++;
  return 0;
}

// exchange one IPC handle per rank through NCCL and open the peers' allocations
static int p2p_open_all(r2s_ctx *ctx, void *local, void **peer_out /*[nranks]*/, const int *want /*ranks to open, -1 terminated*/) {
  // the all-gather below is collective: a rank whose own step fails still takes part and reports the failure afterwards
  cudaIpcMemHandle_t h; memset(&h, 0, sizeof(h));
  bool bad = cudaIpcGetMemHandle(&h, local) != cudaSuccess;
  const int R = ctx->nranks; const size_t W = sizeof(h) / 4;      // 64 bytes = 16 words
  DevBuf tmp; CK(tmp.reserve(sizeof(h) * (size_t)R));
  CK(cudaMemcpyAsync((char *)tmp.p + sizeof(h) * ctx->rank, &h, sizeof(h), cudaMemcpyHostToDevice, ctx->stream));
  NCK(g_nccl.AllGather((char *)tmp.p + sizeof(h) * ctx->rank, tmp.p, W, NC_UINT32, (nccl_comm)ctx->comm, ctx->stream));
  std::vector<cudaIpcMemHandle_t> all((size_t)R);
  if (r2s_readback(ctx, all.data(), tmp.p, sizeof(h) * (size_t)R)) return 1;
  tmp.release();
  for (int i = 0; want[i] >= 0 && !bad; i++) {
    int r = want[i];
    if (r == ctx->rank) { peer_out[r] = local; continue; }
    void *ptr = nullptr;
    if (cudaIpcOpenMemHandle(&ptr, all[(size_t)r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { bad = true; break; }
    peer_out[r] = ptr;
  }
  if (bad) { cudaGetLastError(); FAIL("CUDA IPC mapping of a peer buffer failed"); }
  return 0;
}
// called at the end of r2s_comm_init: returns 0 also when the fast path stays off (then ctx->p2p == false)
static int p2p_setup(r2s_ctx *ctx) {
  ctx->p2p = false;
  if (!ctx->knobs.p2p) return 0;      // R2S_P2P=0: NCCL only
  if (ctx->nranks < 2 || ctx->nranks > 64) return 0;
  // every rank must be able to reach every peer; all ranks take the same decision through a NCCL max-reduce of "cannot"
  int ndev = 0; cudaGetDeviceCount(&ndev);
  CK(cudaMalloc(&ctx->p2p_box, sizeof(P2PBox) + 64));
  CK(cudaMemsetAsync(ctx->p2p_box, 0, sizeof(P2PBox) + 64, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  int want[65]; for (int r = 0; r < ctx->nranks; r++) want[r] = r; want[ctx->nranks] = -1;
  std::string saved = ctx->err;
  int rc = p2p_open_all(ctx, ctx->p2p_box, ctx->p2p_peer_box, want);
  // agree on the outcome
  unsigned *flag = (unsigned *)((char *)ctx->p2p_box + sizeof(P2PBox) + 32), hflag = rc ? 1u : 0u;
  CK(cudaMemcpyAsync(flag, &hflag, sizeof(unsigned), cudaMemcpyHostToDevice, ctx->stream));
  NCK(g_nccl.AllReduce(flag, flag, 1, NC_UINT32, NC_MAX, (nccl_comm)ctx->comm, ctx->stream));
  if (r2s_readback(ctx, &hflag, flag, sizeof(unsigned))) return 1;
  if (hflag) { ctx->err = saved; return 0; }      // some rank could not map its peers: everybody stays on NCCL
  CK(cudaMalloc((void **)&ctx->p2p_peer_box_dev, sizeof(void *) * 64));
  CK(cudaMemcpyAsync(ctx->p2p_peer_box_dev, ctx->p2p_peer_box, sizeof(void *) * (size_t)ctx->nranks, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->p2p = true; ctx->p2p_seq = 0; ctx->p2p_halo_seq = 0;
  return 0;
}
static void p2p_teardown(r2s_ctx *ctx) {
  if (!ctx->p2p_box) return;
  for (int s = 0; s < 2; s++) if (ctx->p2p_c_peer[s]) { cudaIpcCloseMemHandle(ctx->p2p_c_peer[s]); ctx->p2p_c_peer[s] = nullptr; }
  if (ctx->p2p) for (int r = 0; r < ctx->nranks; r++) if (r != ctx->rank && ctx->p2p_peer_box[r]) cudaIpcCloseMemHandle(ctx->p2p_peer_box[r]);
  if (ctx->p2p_peer_box_dev) cudaFree(ctx->p2p_peer_box_dev);
  cudaFree(ctx->p2p_box);
  ctx->p2p_box = nullptr; ctx->p2p_peer_box_dev = nullptr; ctx->p2p = false; ctx->p2p_c_local = nullptr;
}

// ---- in-process group set-up ---------------------------------------------------------------------------------------------
LocalGroup *r2s_local_group_create(r2s_ctx **ctxs, int n, std::string *err) {
  auto fail = [&](const char *m) { if (err) *err = m; return (LocalGroup *)nullptr; };
  if (n < 2 || n > 64) return fail("in-process slab group: 2..64 slabs");
  LocalGroup *lg = new LocalGroup();
  lg->n = n;
  for (int r = 0; r < n; r++) { lg->ctx[r] = ctxs[r]; lg->ev_ready[r] = nullptr; lg->ev_done[r] = nullptr; lg->stage[r] = nullptr; }
  lg->spin_ok = true;
  for (int r = 0; r < n; r++) for (int q = 0; q < r; q++) if (ctxs[q]->device == ctxs[r]->device) lg->spin_ok = false;
  // peer access between every pair of distinct devices (slabs that share a device need none)
  for (int r = 0; r < n; r++) {
    if (cudaSetDevice(ctxs[r]->device) != cudaSuccess) { delete lg; return fail("cudaSetDevice failed"); }
    for (int q = 0; q < n; q++) {
      if (ctxs[q]->device == ctxs[r]->device) continue;
      int can = 0; cudaDeviceCanAccessPeer(&can, ctxs[r]->device, ctxs[q]->device);
      if (!can) { delete lg; return fail("in-process slab group: the devices cannot access each other's memory (no NVLink / PCIe peer access)"); }
      cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[q]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { delete lg; return fail("cudaDeviceEnablePeerAccess failed"); }
      cudaGetLastError();
    }
    if (cudaEventCreateWithFlags(&lg->ev_ready[r], cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&lg->ev_done[r], cudaEventDisableTiming) != cudaSuccess) { delete lg; return fail("cudaEventCreate failed"); }
    if (cudaMalloc(&ctxs[r]->p2p_box, sizeof(P2PBox) + 64) != cudaSuccess || cudaMemset(ctxs[r]->p2p_box, 0, sizeof(P2PBox) + 64) != cudaSuccess) { delete lg; return fail("mailbox allocation failed"); }
    if (cudaMalloc(&lg->stage[r], 64 * 32) != cudaSuccess) { delete lg; return fail("staging allocation failed"); }
  }
  for (int r = 0; r < n; r++) {
    r2s_ctx *c = ctxs[r];
    cudaSetDevice(c->device);
    for (int q = 0; q < n; q++) c->p2p_peer_box[q] = ctxs[q]->p2p_box;
    if (cudaMalloc((void **)&c->p2p_peer_box_dev, sizeof(void *) * 64) != cudaSuccess ||
        cudaMemcpy(c->p2p_peer_box_dev, c->p2p_peer_box, sizeof(void *) * (size_t)n, cudaMemcpyHostToDevice) != cudaSuccess) { delete lg; return fail("mailbox table allocation failed"); }
    c->lg = lg; c->rank = r; c->nranks = n; c->comm = nullptr; c->p2p = lg->spin_ok; c->p2p_seq = 0; c->p2p_halo_seq = 0; c->p2p_c_local = nullptr;
  }
  return lg;
}
void r2s_local_group_destroy(LocalGroup *lg) {
  if (!lg) return;
  for (int r = 0; r < lg->n; r++) {
    if (lg->ctx[r]) { cudaSetDevice(lg->ctx[r]->device); r2s_comm_destroy(lg->ctx[r]); }
    if (lg->ev_ready[r]) cudaEventDestroy(lg->ev_ready[r]);
    if (lg->ev_done[r]) cudaEventDestroy(lg->ev_done[r]);
    if (lg->stage[r]) cudaFree(lg->stage[r]);
  }
  delete lg;
}

// ---- CG halo planes of c over peer memory ------------------------------------------------------------------------------
// Collective, called BEFORE the CG vector c may be re-allocated (a larger grid on a live communicator): if any rank is going to grow
// its buffer, every rank first closes its CUDA-IPC mappings of the neighbours' c (freeing memory that a peer still has mapped is
// undefined behaviour, and a re-allocation that happens to return the old address would otherwise leave a dead mapping in use);
// a second all-reduce makes sure everybody has closed before anybody frees.  r2s_p2p_map_c then maps the new buffers.
int r2s_p2p_prepare_c(r2s_ctx *ctx, size_t have_bytes, size_t need_bytes) {
  if (!ctx->p2p || ctx->lg) return 0;      // in-process groups use plain pointers, re-read on every call
  unsigned long long *w = (unsigned long long *)((char *)ctx->p2p_box + sizeof(P2PBox) + 40);
  unsigned long long grow = have_bytes < need_bytes ? 1ull : 0ull;
  CK(cudaMemcpyAsync(w, &grow, sizeof(grow), cudaMemcpyHostToDevice, ctx->stream));
  if (p2p_allreduce(ctx, w, 1, 4)) return 1;
  if (r2s_readback(ctx, &grow, w, sizeof(grow))) return 1;
  if (!grow) return 0;
  for (int s = 0; s < 2; s++) if (ctx->p2p_c_peer[s]) { cudaIpcCloseMemHandle(ctx->p2p_c_peer[s]); ctx->p2p_c_peer[s] = nullptr; }
  ctx->p2p_c_local = nullptr;      // forces the re-mapping in r2s_p2p_map_c on every rank
  if (p2p_allreduce(ctx, w, 1, 4)) return 1;
  CK(cudaStreamSynchronize(ctx->stream));
  return 0;
}
int r2s_p2p_map_c(r2s_ctx *ctx, float *c, size_t bytes) {
  if (!ctx->p2p) return 0;
  if (ctx->lg) {      // in-process: the neighbours' arrays are plain pointers, re-read on every call (they may have been re-allocated)
    LocalGroup *lg = ctx->lg;
    lg->src[ctx->rank] = c;
    LBAR(lg);
    ctx->p2p_c_peer[0] = ctx->rank > 0 ? (void *)lg->src[ctx->rank - 1] : nullptr;
    ctx->p2p_c_peer[1] = ctx->rank + 1 < ctx->nranks ? (void *)lg->src[ctx->rank + 1] : nullptr;
    ctx->p2p_c_local = c;
    LBAR(lg);      // nobody overwrites src[] before everybody has read it
    return 0;
  }
  // collective decision: remap if any rank's buffer moved
  unsigned long long *w = (unsigned long long *)((char *)ctx->p2p_box + sizeof(P2PBox) + 40);
  unsigned long long changed = (ctx->p2p_c_local != (void *)c) ? 1ull : 0ull;
  CK(cudaMemcpyAsync(w, &changed, sizeof(changed), cudaMemcpyHostToDevice, ctx->stream));
  if (p2p_allreduce(ctx, w, 1, 4)) return 1;
  if (r2s_readback(ctx, &changed, w, sizeof(changed))) return 1;
  if (!changed) return 0;
  for (int s = 0; s < 2; s++) if (ctx->p2p_c_peer[s]) { cudaIpcCloseMemHandle(ctx->p2p_c_peer[s]); ctx->p2p_c_peer[s] = nullptr; }
  void *peers[64]; memset(peers, 0, sizeof(peers));
  int want[3], nw = 0;
  if (ctx->rank > 0) want[nw++] = ctx->rank - 1;
  if (ctx->rank + 1 < ctx->nranks) want[nw++] = ctx->rank + 1;
  want[nw] = -1;
  if (p2p_open_all(ctx, c, peers, want)) return 1;
  ctx->p2p_c_peer[0] = ctx->rank > 0 ? peers[ctx->rank - 1] : nullptr;
  ctx->p2p_c_peer[1] = ctx->rank + 1 < ctx->nranks ? peers[ctx->rank + 1] : nullptr;
  ctx->p2p_c_local = c;
  (void)bytes;
  return 0;
}
// parameters of the fused CG kernels for one iteration (r2s_rbf.cu): consumes two all-reduce sequence numbers and one halo sequence number
int r2s_p2p_fuse_params(r2s_ctx *ctx, P2PFuse *fs, P2PFuse *fu, i64 plane_elems, int k0, int k1, int H) {
  P2PFuse f; memset(&f, 0, sizeof(f));
  if (!ctx->p2p) { *fs = f; *fu = f; return 0; }
  const int r = ctx->rank; const bool lo = r > 0, hi = r + 1 < ctx->nranks; const int h = std::min(H, k1 - k0);
  f.enabled = 1; f.rank = r; f.R = ctx->nranks;
  f.mine = (P2PBox *)ctx->p2p_box; f.peers = (P2PBox *const *)ctx->p2p_peer_box_dev;
  f.box_lower = lo ? (P2PBox *)ctx->p2p_peer_box[r - 1] : nullptr; f.box_upper = hi ? (P2PBox *)ctx->p2p_peer_box[r + 1] : nullptr;
  f.c_lower = lo ? (float *)ctx->p2p_c_peer[0] : nullptr; f.c_upper = hi ? (float *)ctx->p2p_c_peer[1] : nullptr;
  f.lo0 = lo ? (i64)k0 * plane_elems : 0; f.lo1 = lo ? (i64)(k0 + h) * plane_elems : 0;
  f.hi0 = hi ? (i64)(k1 - h) * plane_elems : 0; f.hi1 = hi ? (i64)k1 * plane_elems : 0;
  f.seq_halo = ++ctx->p2p_halo_seq;
  f.seq_ar = ++ctx->p2p_seq; *fs = f;
  f.seq_ar = ++ctx->p2p_seq; *fu = f;
  ctx->p2p_ops += 3;
  return 0;
}
// error flag of the mailbox (bounded spins): checked by the pipeline after its final synchronisation
int r2s_p2p_check(r2s_ctx *ctx) {
  if (!ctx->p2p) return 0;
  unsigned long long e = 0;
  if (r2s_readback(ctx, &e, &((P2PBox *)ctx->p2p_box)->error, sizeof(e))) return 1;
  if (e) FAIL("peer-memory exchange timed out (a rank did not arrive; the communicator is unusable afterwards); set R2S_P2P=0 to use NCCL only");
  return 0;
}
