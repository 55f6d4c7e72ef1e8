// exact_host.cpp -- the decision-critical device functions of rho2sdf.jl_b200/csrc/r2s_exact.cuh compiled for the HOST (g++ with
// -ffp-contract=off: the round-to-nearest intrinsics become the plain IEEE operations they stand for), so that their bit-for-bit
// agreement with the CPU oracle can be fuzzed without a GPU.  Test infrastructure only.
#include <math.h>
#ifndef __device__
#define __device__
#endif
#ifndef __forceinline__
#define __forceinline__ inline __attribute__((always_inline))
#endif
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline double __dsqrt_rn(double a) { return sqrt(a); }
#include "../../rho2sdf.jl_b200/csrc/r2s_exact.cuh"

extern "C" {
// Xe: [8][3] row-major nodes; returns 1 on success, xi[3] (10,10,10 on failure)
int exact_host_inverse_map_hex8(const double *Xe, const double *x, double *xi) {
  double T[3][8];
  for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) T[d][a] = Xe[3 * a + d];
  return ex::inverse_map_hex8(T, x, xi) ? 1 : 0;
}
// the per-element affine path of the sign kernel: prepare + apply (returns -1 when the element is not affine)
int exact_host_inverse_map_affine(const double *Xe, const double *x, double *xi) {
  double T[3][8], A[3][8];
  for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) T[d][a] = Xe[3 * a + d];
  for (int d = 0; d < 3; d++) ex::mono8(T[d], A[d]);
  ex::AffineInv S; ex::affine_inverse_prepare(A, S);
  if (!S.affine) return -1;
  return ex::affine_inverse_apply(S, x, xi) ? 1 : 0;
}
int exact_host_cell_range_axis(double lo, double hi, double delta, double amin, double amax, int N, int *I0, int *I1) {
  return ex::cell_range_axis(lo, hi, delta, amin, amax, N, *I0, *I1) ? 1 : 0;
}
void exact_host_barycentric(const double *x1, const double *x2, const double *x3, const double *n, const double *x, double *lam) { ex::barycentric(x1, x2, x3, n, x, lam); }
}
