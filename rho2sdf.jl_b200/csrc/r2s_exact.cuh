// r2s_exact.cuh -- decision-critical double arithmetic, evaluated in a FIXED operation order with
// round-to-nearest intrinsics so the compiler can never contract a multiply and an add into an FMA.
// Every expression here follows the reference formula it cites (paths under src/ of kopacja/rho2sdf.jl)
// in the order Julia evaluates it; a branch taken on one of these values is therefore reproducible
// bit-for-bit on any IEEE-754 machine.
#pragma once
#include <cuda_runtime.h>

namespace ex {
__device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dvd(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double sqr(double a) { return __dsqrt_rn(a); }
__device__ __forceinline__ double norm3(double a, double b, double c) { return sqr(add(add(mul(a, a), mul(b, b)), mul(c, c))); }
__device__ __forceinline__ double max3abs(double a, double b, double c) { return fmax(fabs(a), fmax(fabs(b), fabs(c))); }

// ShapeFunctions/hex8_shape.jl:73-108
__device__ __forceinline__ void hex8_shape(const double xi[3], double N[8]) {
  double m1 = sub(xi[0], 1.0), p1 = add(xi[0], 1.0), m2 = sub(xi[1], 1.0), p2 = add(xi[1], 1.0), m3 = sub(xi[2], 1.0), p3 = add(xi[2], 1.0);
  double t1 = mul(m1, m2), t2 = mul(p1, m2), t3 = mul(p1, p2), t4 = mul(m1, p2);
  const double c = 0.125;
  N[0] = mul(mul(-c, t1), m3); N[1] = mul(mul(c, t2), m3); N[2] = mul(mul(-c, t3), m3); N[3] = mul(mul(c, t4), m3);
  N[4] = mul(mul(c, t1), p3);  N[5] = mul(mul(-c, t2), p3); N[6] = mul(mul(c, t3), p3);  N[7] = mul(mul(-c, t4), p3);
}
// ShapeFunctions/hex8_shape.jl:2-70
__device__ __forceinline__ void hex8_shape_d(const double xi[3], double N[8], double dN[8][3]) {
  double m1 = sub(xi[0], 1.0), p1 = add(xi[0], 1.0), m2 = sub(xi[1], 1.0), p2 = add(xi[1], 1.0), m3 = sub(xi[2], 1.0), p3 = add(xi[2], 1.0);
  double t1 = mul(m1, m2), t2 = mul(p1, m2), t3 = mul(p1, p2), t4 = mul(m1, p2);
  const double c = 0.125;
  N[0] = mul(mul(-c, t1), m3); N[1] = mul(mul(c, t2), m3); N[2] = mul(mul(-c, t3), m3); N[3] = mul(mul(c, t4), m3);
  N[4] = mul(mul(c, t1), p3);  N[5] = mul(mul(-c, t2), p3); N[6] = mul(mul(c, t3), p3);  N[7] = mul(mul(-c, t4), p3);
  double d = mul(c, m3), dp = mul(c, p3);
  dN[0][0] = mul(-d, m2); dN[1][0] = mul(d, m2); dN[2][0] = mul(-d, p2); dN[3][0] = mul(d, p2);
  dN[4][0] = mul(dp, m2); dN[5][0] = mul(-dp, m2); dN[6][0] = mul(dp, p2); dN[7][0] = mul(-dp, p2);
  dN[0][1] = mul(-d, m1); dN[1][1] = mul(d, p1); dN[2][1] = mul(-d, p1); dN[3][1] = mul(d, m1);
  dN[4][1] = mul(dp, m1); dN[5][1] = mul(-dp, p1); dN[6][1] = mul(dp, p1); dN[7][1] = mul(-dp, m1);
  dN[0][2] = mul(-c, t1); dN[1][2] = mul(c, t2); dN[2][2] = mul(-c, t3); dN[3][2] = mul(c, t4);
  dN[4][2] = mul(c, t1);  dN[5][2] = mul(-c, t2); dN[6][2] = mul(c, t3);  dN[7][2] = mul(-c, t4);
}
__device__ __forceinline__ double dot8(const double a[8], const double b[8]) {
  double s = mul(a[0], b[0]);
#pragma unroll
  for (int k = 1; k < 8; k++) s = add(s, mul(a[k], b[k]));
  return s;
}

// monomial coefficients of a trilinear field v = A0 + A1 x + A2 e + A3 z + A4 xe + A5 ez + A6 zx + A7 xez from its nodal values
__device__ __forceinline__ void mono8(const double v[8], double A[8]) {
  double s01 = add(v[0], v[1]), d01 = sub(v[1], v[0]), s32 = add(v[3], v[2]), d32 = sub(v[2], v[3]);
  double s45 = add(v[4], v[5]), d45 = sub(v[5], v[4]), s76 = add(v[7], v[6]), d76 = sub(v[6], v[7]);
  double b0 = add(s01, s32), b1 = add(d01, d32), b2 = sub(s32, s01), b3 = sub(d32, d01);
  double t0 = add(s45, s76), t1 = add(d45, d76), t2 = sub(s76, s45), t3 = sub(d76, d45);
  A[0] = mul(0.125, add(b0, t0)); A[1] = mul(0.125, add(b1, t1)); A[2] = mul(0.125, add(b2, t2)); A[4] = mul(0.125, add(b3, t3));
  A[3] = mul(0.125, sub(t0, b0)); A[6] = mul(0.125, sub(t1, b1)); A[5] = mul(0.125, sub(t2, b2)); A[7] = mul(0.125, sub(t3, b3));
}
// Inverse isoparametric map of a HEX8 (restates SignedDistances/FindLocalCoordinates.jl:16-107 as a converged Newton
// iteration from xi = 0 on the monomial form).  A[d][8] from mono8 of the nodal coordinates, `affine` = all mixed
// coefficients are exactly zero (parallelepiped: one step is exact).  Returns true on success; (10,10,10) otherwise (:106).
__device__ inline bool inverse_map_hex8_mono(const double A[3][8], bool affine, const double x[3], double xi[3]) {
  xi[0] = xi[1] = xi[2] = 0.0;
  for (int it = 0; it < 50; it++) {
    double X = xi[0], E = xi[1], Z = xi[2], xe = mul(X, E), ez = mul(E, Z), zx = mul(Z, X), xez = mul(xe, Z);
    double r[3], J[3][3];
#pragma unroll
    for (int d = 0; d < 3; d++) {
      const double *a = A[d];
      double val = add(add(add(add(add(add(add(a[0], mul(a[1], X)), mul(a[2], E)), mul(a[3], Z)), mul(a[4], xe)), mul(a[5], ez)), mul(a[6], zx)), mul(a[7], xez));
      r[d] = sub(val, x[d]);
      J[d][0] = add(add(add(a[1], mul(a[4], E)), mul(a[6], Z)), mul(a[7], ez));
      J[d][1] = add(add(add(a[2], mul(a[4], X)), mul(a[5], Z)), mul(a[7], zx));
      J[d][2] = add(add(add(a[3], mul(a[5], E)), mul(a[6], X)), mul(a[7], xe));
    }
    double c00 = sub(mul(J[1][1], J[2][2]), mul(J[1][2], J[2][1])), c01 = sub(mul(J[1][2], J[2][0]), mul(J[1][0], J[2][2])),
           c02 = sub(mul(J[1][0], J[2][1]), mul(J[1][1], J[2][0]));
    double det = add(add(mul(J[0][0], c00), mul(J[0][1], c01)), mul(J[0][2], c02));
    if (!(fabs(det) > 0.0)) break;
    double c10 = sub(mul(J[0][2], J[2][1]), mul(J[0][1], J[2][2])), c11 = sub(mul(J[0][0], J[2][2]), mul(J[0][2], J[2][0])),
           c12 = sub(mul(J[0][1], J[2][0]), mul(J[0][0], J[2][1]));
    double c20 = sub(mul(J[0][1], J[1][2]), mul(J[0][2], J[1][1])), c21 = sub(mul(J[0][2], J[1][0]), mul(J[0][0], J[1][2])),
           c22 = sub(mul(J[0][0], J[1][1]), mul(J[0][1], J[1][0]));
    double d0 = dvd(add(add(mul(c00, r[0]), mul(c10, r[1])), mul(c20, r[2])), det);
    double d1 = dvd(add(add(mul(c01, r[0]), mul(c11, r[1])), mul(c21, r[2])), det);
    double d2 = dvd(add(add(mul(c02, r[0]), mul(c12, r[1])), mul(c22, r[2])), det);
    xi[0] = sub(xi[0], d0); xi[1] = sub(xi[1], d1); xi[2] = sub(xi[2], d2);
    double m = max3abs(d0, d1, d2);
    if (!(m < 1.0e3) || !(max3abs(xi[0], xi[1], xi[2]) < 1.0e3)) break;
    if (affine) return true;
    if (m < 1.0e-13) return true;
  }
  xi[0] = xi[1] = xi[2] = 10.0;
  return false;
}
__device__ inline bool inverse_map_hex8(const double Xe[3][8], const double x[3], double xi[3]) {
  double A[3][8]; bool affine = true;
#pragma unroll
  for (int d = 0; d < 3; d++) { mono8(Xe[d], A[d]); if (A[d][4] != 0.0 || A[d][5] != 0.0 || A[d][6] != 0.0 || A[d][7] != 0.0) affine = false; }
  return inverse_map_hex8_mono(A, affine, x, xi);
}
// Per-element data of the inverse map for AFFINE hexes (parallelepipeds: all mixed monomial coefficients are exactly zero).
// The Newton iteration of inverse_map_hex8_mono from xi = 0 is then exact after one step, and everything in that step except
// the right-hand side depends on the element only: a0 (image of the element centre), the cofactors of J and det J.  prepare()
// evaluates them with the very same operations as the general path, apply() finishes the step for one point: 3 subtractions,
// 9 multiplications, 6 additions, 3 divisions -- bit-identical to inverse_map_hex8 on an affine element.
struct AffineInv { double a0[3]; double c[9]; double det; int affine; int pad; };
__device__ inline void affine_inverse_prepare(const double A[3][8], AffineInv &S) {
  bool affine = true;
#pragma unroll
  for (int d = 0; d < 3; d++) if (A[d][4] != 0.0 || A[d][5] != 0.0 || A[d][6] != 0.0 || A[d][7] != 0.0) affine = false;
  S.affine = affine ? 1 : 0; S.pad = 0;
  double J[3][3];
#pragma unroll
  for (int d = 0; d < 3; d++) { S.a0[d] = A[d][0]; J[d][0] = A[d][1]; J[d][1] = A[d][2]; J[d][2] = A[d][3]; }
  S.c[0] = sub(mul(J[1][1], J[2][2]), mul(J[1][2], J[2][1])); S.c[1] = sub(mul(J[1][2], J[2][0]), mul(J[1][0], J[2][2])); S.c[2] = sub(mul(J[1][0], J[2][1]), mul(J[1][1], J[2][0]));
  S.det = add(add(mul(J[0][0], S.c[0]), mul(J[0][1], S.c[1])), mul(J[0][2], S.c[2]));
  S.c[3] = sub(mul(J[0][2], J[2][1]), mul(J[0][1], J[2][2])); S.c[4] = sub(mul(J[0][0], J[2][2]), mul(J[0][2], J[2][0])); S.c[5] = sub(mul(J[0][1], J[2][0]), mul(J[0][0], J[2][1]));
  S.c[6] = sub(mul(J[0][1], J[1][2]), mul(J[0][2], J[1][1])); S.c[7] = sub(mul(J[0][2], J[1][0]), mul(J[0][0], J[1][2])); S.c[8] = sub(mul(J[0][0], J[1][1]), mul(J[0][1], J[1][0]));
}
__device__ __forceinline__ bool affine_inverse_apply(const AffineInv &S, const double x[3], double xi[3]) {
  if (!(fabs(S.det) > 0.0)) { xi[0] = xi[1] = xi[2] = 10.0; return false; }
  const double r0 = sub(S.a0[0], x[0]), r1 = sub(S.a0[1], x[1]), r2 = sub(S.a0[2], x[2]);
  const double d0 = dvd(add(add(mul(S.c[0], r0), mul(S.c[3], r1)), mul(S.c[6], r2)), S.det);
  const double d1 = dvd(add(add(mul(S.c[1], r0), mul(S.c[4], r1)), mul(S.c[7], r2)), S.det);
  const double d2 = dvd(add(add(mul(S.c[2], r0), mul(S.c[5], r1)), mul(S.c[8], r2)), S.det);
  xi[0] = sub(0.0, d0); xi[1] = sub(0.0, d1); xi[2] = sub(0.0, d2);
  if (!(max3abs(d0, d1, d2) < 1.0e3) || !(max3abs(xi[0], xi[1], xi[2]) < 1.0e3)) { xi[0] = xi[1] = xi[2] = 10.0; return false; }
  return true;
}
// TET4: FindLocalCoordinates.jl:110-149 (adjugate solve), returns validity per ElementTypes.jl:104-106
__device__ inline bool inverse_map_tet4(const double Xe[3][4], const double x[3], double lc[3]) {
  double A[3][3], b[3];
#pragma unroll
  for (int d = 0; d < 3; d++) { A[d][0] = sub(Xe[d][1], Xe[d][0]); A[d][1] = sub(Xe[d][2], Xe[d][0]); A[d][2] = sub(Xe[d][3], Xe[d][0]); b[d] = sub(x[d], Xe[d][0]); }
  double c00 = sub(mul(A[1][1], A[2][2]), mul(A[1][2], A[2][1])), c01 = sub(mul(A[1][2], A[2][0]), mul(A[1][0], A[2][2])),
         c02 = sub(mul(A[1][0], A[2][1]), mul(A[1][1], A[2][0]));
  double det = add(add(mul(A[0][0], c00), mul(A[0][1], c01)), mul(A[0][2], c02));
  if (!(fabs(det) > 0.0)) { lc[0] = lc[1] = lc[2] = 10.0; return false; }
  double c10 = sub(mul(A[0][2], A[2][1]), mul(A[0][1], A[2][2])), c11 = sub(mul(A[0][0], A[2][2]), mul(A[0][2], A[2][0])),
         c12 = sub(mul(A[0][1], A[2][0]), mul(A[0][0], A[2][1]));
  double c20 = sub(mul(A[0][1], A[1][2]), mul(A[0][2], A[1][1])), c21 = sub(mul(A[0][2], A[1][0]), mul(A[0][0], A[1][2])),
         c22 = sub(mul(A[0][0], A[1][1]), mul(A[0][1], A[1][0]));
  double l2 = dvd(add(add(mul(c00, b[0]), mul(c10, b[1])), mul(c20, b[2])), det);
  double l3 = dvd(add(add(mul(c01, b[0]), mul(c11, b[1])), mul(c21, b[2])), det);
  double l4 = dvd(add(add(mul(c02, b[0]), mul(c12, b[1])), mul(c22, b[2])), det);
  double l1 = sub(1.0, add(add(l2, l3), l4));
  if (l1 >= 0.0 && l2 >= 0.0 && l3 >= 0.0 && l4 >= 0.0 && add(add(add(l1, l2), l3), l4) <= 1.0) { lc[0] = l1; lc[1] = l2; lc[2] = l3; return true; }
  lc[0] = lc[1] = lc[2] = 10.0;
  return false;
}

// barycentricCoordinates (SignedDistances/TriangularMeshUtils.jl:1-24); LU with partial pivoting
__device__ inline void barycentric(const double x1[3], const double x2[3], const double x3[3], const double n[3], const double x[3], double lam[3]) {
  double A[3][3], b[3];
  A[0][0] = sub(mul(x1[1], n[2]), mul(x1[2], n[1])); A[0][1] = sub(mul(x2[1], n[2]), mul(x2[2], n[1])); A[0][2] = sub(mul(x3[1], n[2]), mul(x3[2], n[1]));
  A[1][0] = sub(mul(x1[2], n[0]), mul(x1[0], n[2])); A[1][1] = sub(mul(x2[2], n[0]), mul(x2[0], n[2])); A[1][2] = sub(mul(x3[2], n[0]), mul(x3[0], n[2]));
  A[2][0] = sub(mul(x1[0], n[1]), mul(x1[1], n[0])); A[2][1] = sub(mul(x2[0], n[1]), mul(x2[1], n[0])); A[2][2] = sub(mul(x3[0], n[1]), mul(x3[1], n[0]));
  b[0] = sub(mul(x[1], n[2]), mul(x[2], n[1])); b[1] = sub(mul(x[2], n[0]), mul(x[0], n[2])); b[2] = sub(mul(x[0], n[1]), mul(x[1], n[0]));
  int im = 0;
  if (fabs(n[1]) > fabs(n[im])) im = 1;
  if (fabs(n[2]) > fabs(n[im])) im = 2;
#pragma unroll
  for (int r = 0; r < 3; r++)
    if (r == im) { A[r][0] = A[r][1] = A[r][2] = 1.0; b[r] = 1.0; }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    int p = k;
#pragma unroll
    for (int r = k + 1; r < 3; r++)
      if (fabs(A[r][k]) > fabs(A[p][k])) p = r;
    if (p != k) {
#pragma unroll
      for (int r = k + 1; r < 3; r++)
        if (r == p) {
#pragma unroll
          for (int c = 0; c < 3; c++) { double t = A[k][c]; A[k][c] = A[r][c]; A[r][c] = t; }
          double t = b[k]; b[k] = b[r]; b[r] = t;
        }
    }
#pragma unroll
    for (int r = k + 1; r < 3; r++) {
      double l = dvd(A[r][k], A[k][k]);
#pragma unroll
      for (int c = k + 1; c < 3; c++) A[r][c] = sub(A[r][c], mul(l, A[k][c]));
      b[r] = sub(b[r], mul(l, b[k]));
    }
  }
  lam[2] = dvd(b[2], A[2][2]);
  lam[1] = dvd(sub(b[1], mul(A[1][2], lam[2])), A[1][1]);
  lam[0] = dvd(sub(sub(b[0], mul(A[0][1], lam[1])), mul(A[0][2], lam[2])), A[0][0]);
}

// calculateMiniAABB_grid (MeshGrid/Grid.jl:122-154) for one axis: cell range of [lo-delta, hi+delta]; false if empty
__device__ __forceinline__ bool cell_range_axis(double lo, double hi, double delta, double amin, double amax, int N, int &I0, int &I1) {
  double a = sub(lo, delta), b = add(hi, delta);
  double f0 = floor(dvd(mul((double)N, sub(a, amin)), sub(amax, amin)));
  double f1 = floor(dvd(mul((double)N, sub(b, amin)), sub(amax, amin)));
  if (f0 < 0) f0 = 0;
  if (f1 >= (double)N) f1 = (double)N;
  if (!(f0 <= f1)) return false;
  I0 = (int)f0; I1 = (int)f1;
  return true;
}
}  // namespace ex
