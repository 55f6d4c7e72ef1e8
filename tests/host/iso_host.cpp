// iso_host.cpp -- the device functions of rho2sdf.jl_b200/csrc/r2s_iso.cuh compiled for the HOST (g++, no CUDA runtime needed)
// so that the per-lane arithmetic of the projection kernels can be checked against the CPU oracle without a GPU, and so that
// the lane divergence of the warp-level drivers can be simulated offline (tools/divergence_sim.py).  Test infrastructure only.
#include <math.h>
#include <string.h>
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
#define R2S_ISO_HOST 1
#include "../../rho2sdf.jl_b200/csrc/r2s_iso.cuh"

static const double sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
static const int edges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

static void coefficients(const double *Xe /*[8][3]*/, const double *re, double A[4][8]) {
  for (int c = 0; c < 4; c++) {
    double nv[8];
    for (int k = 0; k < 8; k++) nv[k] = c < 3 ? Xe[3 * k + c] : re[k];
    iso::monomial8(nv, A[c]);
  }
}
static double gscale(const double *re, double rho_t) {
  double gs = fabs(rho_t);
  for (int k = 0; k < 8; k++) gs = fmax(gs, fabs(re[k]));
  return fmax(gs, 1.0);
}

extern "C" {
// variant 0: general trilinear path, 1: box path (returns -2 when the element is not a canonical axis-aligned box)
// out: xi[3], dist; trace (optional, length >= 128): per phase-2 iteration the status code.  Returns 1 converged, 0 failed.
int iso_host_project(const double *Xe, const double *re, const double *x, double rho_t, int variant, double *xi, double *dist, int *nit) {
  double A[4][8]; coefficients(Xe, re, A);
  const double gs = gscale(re, rho_t);
  bool ok; double p[3];
  if (variant == 0) {
    iso::HexTri T{(const double(*)[8])A};
    ok = iso::project_hex8(T, re, sg, edges, x, rho_t, gs, xi, *nit);
    iso::eval_pos(T, xi, p);
  } else {
    if (!iso::is_box(A)) return -2;
    iso::HexBox B; iso::make_box(A, B);
    ok = iso::project_hex8(B, re, sg, edges, x, rho_t, gs, xi, *nit);
    iso::eval_pos(B, xi, p);
  }
  const double d0 = x[0] - p[0], d1 = x[1] - p[1], d2 = x[2] - p[2];
  *dist = sqrt(fma(d2, d2, fma(d1, d1, d0 * d0)));
  return ok ? 1 : 0;
}
void iso_host_counts(long *out, int reset) { for (int k = 0; k < 8; k++) { out[k] = iso_counts[k]; if (reset) iso_counts[k] = 0; } }
// event trace (see ISO_TRACE in r2s_iso.cuh): the caller provides the buffer; a 0 is appended by iso_host_project_many after every pair
void iso_host_trace(int *buf, long cap) { iso_trace = buf; iso_trace_cap = cap; iso_trace_n = 0; }
long iso_host_trace_len() { return iso_trace_n; }
// TET4: closest point on the in-element iso-polygon (project_tet4); Xe [4][3] row-major nodes.  Returns 1 when a projection exists.
int iso_host_project_tet4(const double *Xe, const double *re, const double *x, double rho_t, double *xp) {
  static const int isn[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
  double T[3][4];
  for (int a = 0; a < 4; a++) for (int d = 0; d < 3; d++) T[d][a] = Xe[3 * a + d];
  return iso::project_tet4(T, re, isn, x, rho_t, xp) ? 1 : 0;
}
int iso_host_is_box(const double *Xe, const double *re) { double A[4][8]; coefficients(Xe, re, A); return iso::is_box(A) ? 1 : 0; }
// batch over n points of ONE element; its[n] receives the phase-2 iteration count (the quantity a warp waits on).
// variant 0: general trilinear, 1: HexBox, 2: HexBox FAST, 3: HexBox FAST with phase 1 computed once for the element (the table
// path of the kernels), 4: general trilinear with phase 1 once per element, 5: as 3 with the single-path tangent step (MODE 3),
// 8: general trilinear with MODE 3 and element phase 1
int iso_host_project_many(const double *Xe, const double *re, long n, const double *x, double rho_t, int variant, double *dist, int *its) {
  double A[4][8]; coefficients(Xe, re, A);
  const double gs = gscale(re, rho_t);
  iso::HexTri T{(const double(*)[8])A}; iso::HexBox B;
  if ((variant >= 1 && variant <= 3) || variant == 5) { if (!iso::is_box(A)) return -2; iso::make_box(A, B); }
  iso::ProjState S0; bool ok0 = false;
  if (variant == 3 || variant == 5) ok0 = iso::proj_init_element<iso::HexBox, 1>(B, rho_t, gs, S0);
  if (variant == 4) ok0 = iso::proj_init_element<iso::HexTri, 0>(T, rho_t, gs, S0);
  if (variant == 8) ok0 = iso::proj_init_element<iso::HexTri, 1>(T, rho_t, gs, S0);
  int bad = 0;
  for (long q = 0; q < n; q++) {
    double xi[3], p[3]; int nit = 0; bool ok;
    const double *xq = x + 3 * q;
    if (variant == 0) { ok = iso::project_hex8(T, re, sg, edges, xq, rho_t, gs, xi, nit); iso::eval_pos(T, xi, p); }
    else if (variant == 1) { ok = iso::project_hex8(B, re, sg, edges, xq, rho_t, gs, xi, nit); iso::eval_pos(B, xi, p); }
    else if (variant == 2) { ok = iso::project_hex8<iso::HexBox, 1>(B, re, sg, edges, xq, rho_t, gs, xi, nit); iso::eval_pos(B, xi, p); }
    else if (variant == 3) { ok = iso::project_hex8_from<iso::HexBox, 1>(B, re, sg, edges, xq, rho_t, gs, S0.xi, ok0, xi, nit); iso::eval_pos(B, xi, p); }
    else if (variant == 5) { ok = iso::project_hex8_from<iso::HexBox, 3>(B, re, sg, edges, xq, rho_t, gs, S0.xi, ok0, xi, nit); iso::eval_pos(B, xi, p); }
    else if (variant == 8) { ok = iso::project_hex8_from<iso::HexTri, 3>(T, re, sg, edges, xq, rho_t, gs, S0.xi, ok0, xi, nit); iso::eval_pos(T, xi, p); }
    else { ok = iso::project_hex8_from<iso::HexTri, 0>(T, re, sg, edges, xq, rho_t, gs, S0.xi, ok0, xi, nit); iso::eval_pos(T, xi, p); }
    const double d0 = xq[0] - p[0], d1 = xq[1] - p[1], d2 = xq[2] - p[2];
    dist[q] = sqrt(fma(d2, d2, fma(d1, d1, d0 * d0))); its[q] = nit; bad += ok ? 0 : 1;
    ISO_TRACE(0);
  }
  return bad;
}
}
