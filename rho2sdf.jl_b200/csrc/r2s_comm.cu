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
#include <string.h>
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
  return 0;
}
extern "C" int r2s_comm_destroy(r2s_ctx *ctx) {
  if (!ctx) return 1;
  if (ctx->comm && g_nccl.lib) { cudaStreamSynchronize(ctx->stream); g_nccl.CommDestroy((nccl_comm)ctx->comm); }
  ctx->comm = nullptr; ctx->rank = 0; ctx->nranks = 1;
  return 0;
}

// fuse the collectives issued between the two calls into one NCCL launch (nested groups are allowed)
int r2s_group_start(r2s_ctx *ctx) { if (ctx->nranks > 1) NCK(g_nccl.GroupStart()); return 0; }
int r2s_group_end(r2s_ctx *ctx) { if (ctx->nranks > 1) NCK(g_nccl.GroupEnd()); return 0; }
// ---- collectives used by the pipeline; all are no-ops for a single rank ----------------------------------------------
int r2s_allreduce(r2s_ctx *ctx, void *buf, size_t count, int kind /*0 f64 sum, 1 u64 sum, 2 u32 max, 3 u32 min, 4 u64 max*/) {
  if (ctx->nranks <= 1) return 0;
  int dt = kind == 0 ? NC_F64 : (kind == 1 || kind == 4 ? NC_UINT64 : NC_UINT32);
  int op = (kind == 0 || kind == 1) ? NC_SUM : (kind == 3 ? NC_MIN : NC_MAX);
  NCK(g_nccl.AllReduce(buf, buf, count, dt, op, (nccl_comm)ctx->comm, ctx->stream));
  ctx->collectives++;
  return 0;
}
int r2s_allgather_u32(r2s_ctx *ctx, const unsigned *send, unsigned *recv, size_t count) {
  if (ctx->nranks <= 1) { CK(cudaMemcpyAsync(recv, send, count * sizeof(unsigned), cudaMemcpyDeviceToDevice, ctx->stream)); return 0; }
  NCK(g_nccl.AllGather(send, recv, count, NC_UINT32, (nccl_comm)ctx->comm, ctx->stream));
  ctx->collectives++;
  return 0;
}
// Exchange halo planes of a globally indexed field a[plane * plane_elems ...] (4-byte elements): this rank owns planes
// [k0, k1) of nz (coarse or fine planes); after the call planes [k0 - below, k0) and [k1, k1 + above) hold the neighbours' values (clipped to [0, nz)).
int r2s_halo_exchange_f32(r2s_ctx *ctx, float *a, i64 plane_elems, int k0, int k1, int nz, int below, int above) {
  if (ctx->nranks <= 1) return 0;
  const int r = ctx->rank;
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
