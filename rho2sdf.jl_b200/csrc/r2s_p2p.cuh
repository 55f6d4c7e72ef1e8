// r2s_p2p.cuh -- the peer-memory mailbox (r2s_comm.cu) and the device-side pieces that the CG kernels of r2s_rbf.cu use to do their
// exchanges THEMSELVES: remote stores of the halo planes while computing, an all-reduce through the mailbox by the last CTA.
#pragma once
#include <cuda_runtime.h>
#define P2P_SLOT_WORDS 8
#define P2P_MAX_SPIN (1u << 26)      // bounded waits (tens of seconds): a dead peer raises the error flag instead of hanging the GPU
struct P2PBox {
  unsigned long long slot[2][64][P2P_SLOT_WORDS];
  unsigned long long halo_flag[2][2];      // [parity][0 = from lower neighbour, 1 = from upper neighbour]
  unsigned long long error;
  unsigned long long halo_count[2];        // CTA completion counters of the put kernel
};

// What a fused compute+exchange kernel needs (passed by value; enabled = 0: single GPU or NCCL / event transport, nothing happens)
struct P2PFuse {
  int enabled, rank, R;
  unsigned seq_ar, seq_halo;                 // sequence numbers of this kernel's all-reduce and of the halo exchange of this iteration
  P2PBox *mine; P2PBox *const *peers;        // my mailbox, table of all mailboxes (peer-mapped)
  P2PBox *box_lower, *box_upper;             // neighbours' mailboxes (halo flags) or nullptr
  float *c_lower, *c_upper;                  // neighbours' CG vector c (peer-mapped, same global indexing) or nullptr
  long long lo0, lo1, hi0, hi1;              // element ranges of c that are the lower / upper neighbour's halo planes
};
// All-reduce (sum in rank order: deterministic, identical on every rank) of one double by ONE CTA (>= 64 threads, all of them call):
// thread p stores this rank's value into slot[rank] of peer p's mailbox, fences and stores the sequence number; then waits until the
// local slot[p] carries it.  Bounded spins: a timeout raises the mailbox's error word.
__device__ __forceinline__ double p2p_allreduce_cta(const P2PFuse &F, unsigned seq, double val) {
  __shared__ double s_tot;
  const int p = threadIdx.x, par = seq & 1;
  if (p < F.R) {
    volatile unsigned long long *dst = F.peers[p]->slot[par][F.rank];
    dst[0] = (unsigned long long)__double_as_longlong(val);
    __threadfence_system();
    dst[P2P_SLOT_WORDS - 1] = (unsigned long long)seq;
    volatile unsigned long long *src = F.mine->slot[par][p];
    unsigned spin = 0;
    const unsigned lim = F.mine->error ? 1024u : P2P_MAX_SPIN;      // after a first time-out the later waits give up quickly: the call is failing anyway
    while (src[P2P_SLOT_WORDS - 1] != (unsigned long long)seq) { if (++spin > lim) { F.mine->error = 1; break; } }
  }
  __threadfence_system();
  __syncthreads();
  if (p == 0) {
    double acc = __longlong_as_double((long long)((volatile unsigned long long *)F.mine->slot[par][0])[0]);
    for (int q = 1; q < F.R; q++) acc += __longlong_as_double((long long)((volatile unsigned long long *)F.mine->slot[par][q])[0]);
    s_tot = acc;
  }
  __syncthreads();
  return s_tot;
}
// thread 0: tell the neighbours that every halo value of this iteration has been stored into their arrays (call after a system fence
// that follows the completion of ALL CTAs' stores)
__device__ __forceinline__ void p2p_raise_halo_flags(const P2PFuse &F) {
  const int par = F.seq_halo & 1;
  __threadfence_system();
  if (F.box_lower) ((volatile unsigned long long *)F.box_lower->halo_flag[par])[1] = F.seq_halo;      // I am the lower neighbour's UPPER neighbour
  if (F.box_upper) ((volatile unsigned long long *)F.box_upper->halo_flag[par])[0] = F.seq_halo;
}
// thread 0: wait for both neighbours' flags of this iteration
__device__ __forceinline__ void p2p_wait_halo_flags(const P2PFuse &F) {
  const int par = F.seq_halo & 1;
  for (int side = 0; side < 2; side++) {
    if (!(side == 0 ? F.box_lower : F.box_upper)) continue;
    volatile unsigned long long *f = &F.mine->halo_flag[par][side];
    unsigned spin = 0;
    const unsigned lim = F.mine->error ? 1024u : P2P_MAX_SPIN;
    while (*f != (unsigned long long)F.seq_halo) { if (++spin > lim) { F.mine->error = 1; break; } }
  }
  __threadfence_system();
}
