/*
 * r2s_oracle.c -- CPU ORACLE for the rho2sdf grid-sampling hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load it.  The product path (the CUDA library
 * under rho2sdf.jl_b200/csrc) never links, imports or calls anything in this file.
 *
 * It is a plain-C restatement of the algorithms of kopacja/rho2sdf.jl (all citations are file:line
 * under /root/reference/src).  The reference delegates two sub-problems to NLopt 2.10.0 (NLopt_jll
 * 2.10.0+0, NLopt.jl 1.2.1 -- Manifest.toml:1112-1126), whose source is not in the reference tree:
 *   - closest point on the in-element iso-surface   (SignedDistances/ComputeCoordsOnIso.jl:16-87, :LD_SLSQP)
 *   - inverse isoparametric map                      (SignedDistances/FindLocalCoordinates.jl:16-107, :LD_LBFGS)
 * They are restated here as *tightly converged* local solvers of the same mathematical problems
 * started from the same start point (xi = 0): a feasible-path SQP (exact Lagrangian Hessian, active
 * set on the box bounds) and a Newton inverse map.  See DESIGN.md "Oracle" for what is and is not
 * pinned by the reference's golden values.
 *
 * Decision-critical arithmetic (cell binning, barycentric coordinates, edge tests, inverse map,
 * sign tests) is written as straight-line IEEE double arithmetic in a fixed order and must be built
 * WITHOUT fused-multiply-add contraction (-ffp-contract=off); the CUDA kernels evaluate the very
 * same expressions with __dmul_rn/__dadd_rn so that every branch decision is bit-identical.
 *
 * Array conventions = Julia's: column-major, X[3*n+d], IEN[nen*e+a] holding 1-based node ids,
 * grid point linear id = k*(N1+1)*(N2+1) + j*(N1+1) + i (Grid.jl:84-90).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;
#define API __attribute__((visibility("default")))

#define BIG 1.0e10

/* ------------------------------------------------------------------------------------------------
 * Element topology tables (ElementTypes/ElementTypes.jl:15-78), 0-based local node ids
 * ---------------------------------------------------------------------------------------------- */
static const int HEX_ISN[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
static const int TET_ISN[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
static const int HEX_EDGES[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};
/* natural coordinates of the HEX8 nodes implied by hex8_shape.jl:27-34 */
static const double HEX_SG[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};

/* ------------------------------------------------------------------------------------------------
 * Shape functions (ShapeFunctions/hex8_shape.jl:2-70 and :73-108), same operation order
 * ---------------------------------------------------------------------------------------------- */
static void hex8_shape(const double xi[3], double N[8]) {
  double m1 = xi[0] - 1, p1 = xi[0] + 1, m2 = xi[1] - 1, p2 = xi[1] + 1, m3 = xi[2] - 1, p3 = xi[2] + 1;
  double t1 = m1 * m2, t2 = p1 * m2, t3 = p1 * p2, t4 = m1 * p2, c = 0.125;
  N[0] = -c * t1 * m3; N[1] = c * t2 * m3; N[2] = -c * t3 * m3; N[3] = c * t4 * m3;
  N[4] = c * t1 * p3;  N[5] = -c * t2 * p3; N[6] = c * t3 * p3; N[7] = -c * t4 * p3;
}
static void hex8_shape_d(const double xi[3], double N[8], double dN[8][3]) {
  double m1 = xi[0] - 1, p1 = xi[0] + 1, m2 = xi[1] - 1, p2 = xi[1] + 1, m3 = xi[2] - 1, p3 = xi[2] + 1;
  double t1 = m1 * m2, t2 = p1 * m2, t3 = p1 * p2, t4 = m1 * p2, c = 0.125;
  N[0] = -c * t1 * m3; N[1] = c * t2 * m3; N[2] = -c * t3 * m3; N[3] = c * t4 * m3;
  N[4] = c * t1 * p3;  N[5] = -c * t2 * p3; N[6] = c * t3 * p3; N[7] = -c * t4 * p3;
  double d = c * m3, dp = c * p3;
  dN[0][0] = -d * m2; dN[1][0] = d * m2; dN[2][0] = -d * p2; dN[3][0] = d * p2;
  dN[4][0] = dp * m2; dN[5][0] = -dp * m2; dN[6][0] = dp * p2; dN[7][0] = -dp * p2;
  dN[0][1] = -d * m1; dN[1][1] = d * p1; dN[2][1] = -d * p1; dN[3][1] = d * m1;
  dN[4][1] = dp * m1; dN[5][1] = -dp * p1; dN[6][1] = dp * p1; dN[7][1] = -dp * m1;
  dN[0][2] = -c * t1; dN[1][2] = c * t2; dN[2][2] = -c * t3; dN[3][2] = c * t4;
  dN[4][2] = c * t1;  dN[5][2] = -c * t2; dN[6][2] = c * t3;  dN[7][2] = -c * t4;
}

static double det3(const double J[3][3]) {
  return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}
static inline double norm3(const double v[3]) { return sqrt((v[0] * v[0] + v[1] * v[1]) + v[2] * v[2]); }

/* Gauss-Legendre nodes/weights (FastGaussQuadrature.gausslegendre), Newton on P_n in long double */
static void gauss_legendre(int n, double *x, double *w) {
  for (int i = 0; i < n; i++) {
    long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L)), pp = 0;
    for (int it = 0; it < 100; it++) {
      long double p1 = 1, p2 = 0;
      for (int j = 0; j < n; j++) { long double p3 = p2; p2 = p1; p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1);
      long double dz = p1 / pp; z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    { long double p1 = 1, p2 = 0;
      for (int j = 0; j < n; j++) { long double p3 = p2; p2 = p1; p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1); }
    /* ascending order like FastGaussQuadrature: node i from the left */
    x[n - 1 - i] = (double)z; w[n - 1 - i] = (double)(2 / ((1 - z * z) * pp * pp));
  }
  for (int i = 0; i < n / 2; i++) { /* enforce exact antisymmetry */
    double a = 0.5 * (x[n - 1 - i] - x[i]); x[i] = -a; x[n - 1 - i] = a;
    double b = 0.5 * (w[i] + w[n - 1 - i]); w[i] = b; w[n - 1 - i] = b;
  }
  if (n & 1) x[n / 2] = 0.0;
}

/* ------------------------------------------------------------------------------------------------
 * node -> element connectivity INE (MeshGrid/MeshInformations.jl:69-77), ascending element order
 * ---------------------------------------------------------------------------------------------- */
typedef struct { i64 *ptr; i64 *el; } ine_t;
static ine_t build_ine(i64 nnp, i64 nel, int nen, const i64 *IEN) {
  ine_t t; t.ptr = (i64 *)calloc((size_t)nnp + 1, sizeof(i64));
  for (i64 e = 0; e < nel; e++) for (int a = 0; a < nen; a++) t.ptr[IEN[nen * e + a]]++;   /* 1-based id -> slot id */
  for (i64 n = 0; n < nnp; n++) t.ptr[n + 1] += t.ptr[n];
  t.el = (i64 *)malloc(sizeof(i64) * (size_t)t.ptr[nnp]);
  i64 *fill = (i64 *)calloc((size_t)nnp, sizeof(i64));
  for (i64 e = 0; e < nel; e++) for (int a = 0; a < nen; a++) { i64 n = IEN[nen * e + a] - 1; t.el[t.ptr[n] + fill[n]++] = e; }
  free(fill); return t;
}
static void free_ine(ine_t *t) { free(t->ptr); free(t->el); }

/* ------------------------------------------------------------------------------------------------
 * calculate_mesh_volume (MeshGrid/MeshVolume.jl:4-117)
 * ---------------------------------------------------------------------------------------------- */
static double elem_volume_hex8(const double *X, const i64 *en, const double *gp, const double *w, int n) {
  double xe[3][8];
  for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) xe[d][a] = X[3 * (en[a] - 1) + d];
  double vol = 0.0, N[8], dN[8][3];
  for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    double xi[3] = {gp[i], gp[j], gp[k]}; hex8_shape_d(xi, N, dN);
    double J[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double s = 0; for (int a = 0; a < 8; a++) s += xe[r][a] * dN[a][c]; J[r][c] = s; }
    vol += w[i] * w[j] * w[k] * fabs(det3(J));
  }
  return vol;
}
static double elem_volume_tet4(const double *X, const i64 *en, const double *gp, const double *w, int n) {
  /* MeshVolume.jl:76-117: cube->tet collapsed quadrature; dN of TET4 is constant (ShapeFunctions.jl:39-73) */
  double xe[3][4];
  for (int a = 0; a < 4; a++) for (int d = 0; d < 3; d++) xe[d][a] = X[3 * (en[a] - 1) + d];
  double J[3][3];
  for (int r = 0; r < 3; r++) { J[r][0] = xe[r][0] - xe[r][3]; J[r][1] = xe[r][1] - xe[r][3]; J[r][2] = xe[r][2] - xe[r][3]; }
  double adet = fabs(det3(J)), vol = 0.0;
  for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
    double xi = (gp[i] + 1.0) / 2.0, eta = (gp[j] + 1.0) / 2.0 * (1.0 - xi), zeta = (gp[k] + 1.0) / 2.0 * (1.0 - xi - eta);
    if (xi < 0 || eta < 0 || zeta < 0 || xi + eta + zeta > 1.0) continue;
    double jt = (1.0 - xi) * (1.0 - xi) * (1.0 - xi - eta) / 8.0;
    vol += w[i] * w[j] * w[k] * adet * jt;
  }
  return vol;
}
API int r2so_mesh_volume(i64 nnp, const double *X, i64 nel, int nen, const i64 *IEN, const double *rho, double *V_domain, double *V_frac) {
  (void)nnp; double gp[3], w[3]; gauss_legendre(3, gp, w);
  double dom = 0.0, to = 0.0;
  for (i64 e = 0; e < nel; e++) {
    double v = nen == 8 ? elem_volume_hex8(X, IEN + 8 * e, gp, w, 3) : elem_volume_tet4(X, IEN + 4 * e, gp, w, 3);
    dom += v; to += v * rho[e];
  }
  *V_domain = dom; *V_frac = to / dom; return 0;
}

/* ------------------------------------------------------------------------------------------------
 * DenseInNodes (MeshGrid/NodalDensities.jl:89-218)
 * ---------------------------------------------------------------------------------------------- */
static void jacobi_eig4(double A[4][4], double lam[4], double V[4][4]) {
  /* cyclic Jacobi, converged to round-off; eigenvalues returned ascending like LinearAlgebra.eigen(Symmetric) */
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) V[i][j] = (i == j);
  for (int sweep = 0; sweep < 60; sweep++) {
    double off = 0; for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) off += A[p][q] * A[p][q];
    if (off == 0.0) break;
    for (int p = 0; p < 4; p++) for (int q = p + 1; q < 4; q++) {
      if (A[p][q] == 0.0) continue;
      double theta = (A[q][q] - A[p][p]) / (2.0 * A[p][q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < 4; k++) { double akp = A[k][p], akq = A[k][q]; A[k][p] = c * akp - s * akq; A[k][q] = s * akp + c * akq; }
      for (int k = 0; k < 4; k++) { double apk = A[p][k], aqk = A[q][k]; A[p][k] = c * apk - s * aqk; A[q][k] = s * apk + c * aqk; }
      for (int k = 0; k < 4; k++) { double vkp = V[k][p], vkq = V[k][q]; V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq; }
    }
  }
  for (int i = 0; i < 4; i++) lam[i] = A[i][i];
  for (int i = 0; i < 4; i++) for (int j = i + 1; j < 4; j++) if (lam[j] < lam[i]) {
    double t = lam[i]; lam[i] = lam[j]; lam[j] = t;
    for (int k = 0; k < 4; k++) { double u = V[k][i]; V[k][i] = V[k][j]; V[k][j] = u; }
  }
}
API int r2so_nodal_densities(i64 nnp, const double *X, i64 nel, int nen, const i64 *IEN, const double *rho, double *rho_n) {
  ine_t ine = build_ine(nnp, nel, nen, IEN);
  double *C = (double *)malloc(sizeof(double) * 3 * (size_t)nel);     /* GeometricCentre :73-82 (mean of nodes) */
  for (i64 e = 0; e < nel; e++) for (int d = 0; d < 3; d++) {
    double s = 0; for (int a = 0; a < nen; a++) s += X[3 * (IEN[nen * e + a] - 1) + d];
    C[3 * e + d] = s / nen;
  }
  for (i64 i = 0; i < nnp; i++) {
    i64 n1 = ine.ptr[i + 1] - ine.ptr[i]; const i64 *els = ine.el + ine.ptr[i];
    if (n1 == 0) { rho_n[i] = 0.0; continue; }
    if (n1 == 1) { rho_n[i] = rho[els[0]]; continue; }                  /* :98-99 */
    if (n1 < 4) {                                                        /* FilterForNodalDensity :116-136 */
      double L[3], Lmax = 0;
      for (i64 j = 0; j < n1; j++) { double v[3]; for (int d = 0; d < 3; d++) v[d] = X[3 * i + d] - C[3 * els[j] + d]; L[j] = norm3(v); if (L[j] > Lmax) Lmax = L[j]; }
      Lmax *= 1.2; double dm = 0, de = 0;
      for (i64 j = 0; j < n1; j++) { dm += rho[els[j]] * (1 - L[j] / Lmax); de += (1 - L[j] / Lmax); }
      rho_n[i] = dm / de; continue;
    }
    /* NodalDensityLeastSquares :145-183 */
    double AtA[4][4] = {{0}}, Atb[4] = {0}, bsum = 0;
    for (i64 j = 0; j < n1; j++) {
      double row[4] = {1.0, C[3 * els[j]], C[3 * els[j] + 1], C[3 * els[j] + 2]}, b = rho[els[j]];
      for (int r = 0; r < 4; r++) { for (int c = 0; c < 4; c++) AtA[r][c] += row[r] * row[c]; Atb[r] += row[r] * b; }
      bsum += b;
    }
    double lam[4], V[4][4]; jacobi_eig4(AtA, lam, V);
    /* LamReduction :192-217 */
    double lmax = lam[0], lmin = lam[0];
    for (int k = 1; k < 4; k++) { if (lam[k] > lmax) lmax = lam[k]; if (lam[k] < lmin) lmin = lam[k]; }
    double e1 = fabs(lmax / lmin), e2 = fabs(lmax / lam[1]), e3 = fabs(lmax / lam[2]);
    int poz = -1;                                 /* 0-based index of first kept eigenvalue, -1 = none */
    if (1e7 > e1 && 3e3 > e2) poz = 0;
    else if (1e7 < e1 && 3e3 > e2) poz = 1;
    else if (1e7 < e1 && 3e3 < e2) poz = (3e3 > e3) ? 2 : 3;
    if (poz < 0) { rho_n[i] = bsum / (double)n1; continue; }
    double b1[4], x2[4] = {0, 0, 0, 0}, xx[4];
    for (int k = 0; k < 4; k++) { double s = 0; for (int r = 0; r < 4; r++) s += V[r][k] * Atb[r]; b1[k] = s; }
    for (int k = poz; k < 4; k++) x2[k] = b1[k] / lam[k];
    for (int r = 0; r < 4; r++) { double s = 0; for (int k = 0; k < 4; k++) s += V[r][k] * x2[k]; xx[r] = s; }
    rho_n[i] = 1.0 * xx[0] + X[3 * i] * xx[1] + X[3 * i + 1] * xx[2] + X[3 * i + 2] * xx[3];
  }
  free(C); free_ine(&ine); return 0;
}

/* ------------------------------------------------------------------------------------------------
 * calculate_isocontour_volume / find_threshold_for_volume (MeshGrid/Isocontour_volume.jl:1-154), HEX8 only
 * ---------------------------------------------------------------------------------------------- */
API double r2so_isocontour_volume(i64 nnp, const double *X, i64 nel, const i64 *IEN, const double *rho_n, double thr) {
  (void)nnp; double gd[15], wd[15], gs[3], ws[3]; gauss_legendre(15, gd, wd); gauss_legendre(3, gs, ws);
  double total = 0.0;
  for (i64 e = 0; e < nel; e++) {
    double ev[8], mn = DBL_MAX, mx = -DBL_MAX, xe[3][8];
    for (int a = 0; a < 8; a++) { ev[a] = rho_n[IEN[8 * e + a] - 1]; if (ev[a] < mn) mn = ev[a]; if (ev[a] > mx) mx = ev[a]; }
    if (mx < thr) continue;
    for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) xe[d][a] = X[3 * (IEN[8 * e + a] - 1) + d];
    int chk = !(mn >= thr), n = chk ? 15 : 3; const double *gp = chk ? gd : gs, *w = chk ? wd : ws;
    double vol = 0.0, N[8], dN[8][3];
    for (int k = 0; k < n; k++) for (int j = 0; j < n; j++) for (int i = 0; i < n; i++) {
      double xi[3] = {gp[i], gp[j], gp[k]}; hex8_shape_d(xi, N, dN);
      if (chk) { double v = 0; for (int a = 0; a < 8; a++) v += N[a] * ev[a]; if (v < thr) continue; }
      double J[3][3];
      for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { double s = 0; for (int a = 0; a < 8; a++) s += xe[r][a] * dN[a][c]; J[r][c] = s; }
      vol += w[i] * w[j] * w[k] * fabs(det3(J));
    }
    total += vol;
  }
  return total;
}
API int r2so_find_threshold(i64 nnp, const double *X, i64 nel, const i64 *IEN, const double *rho_n, double target, double tol, int maxit, double *rho_t) {
  double lo = 0.0, hi = 1.0;
  double vmin = r2so_isocontour_volume(nnp, X, nel, IEN, rho_n, hi), vmax = r2so_isocontour_volume(nnp, X, nel, IEN, rho_n, lo);
  if (target > vmax || target < vmin) return 1;                       /* :93-95 error(...) */
  int it = 0; double best = 0.0, best_err = INFINITY;
  while (it < maxit) {
    double th = (lo + hi) / 2, v = r2so_isocontour_volume(nnp, X, nel, IEN, rho_n, th), err = fabs(v - target) / target;
    if (err < best_err) { best = th; best_err = err; }
    if (err < tol) break;
    if (v > target) lo = th; else hi = th;
    it++;
  }
  *rho_t = best; return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Grid helpers (MeshGrid/Grid.jl:47-68, :81-93, :122-154) -- decision-critical, fixed operation order
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  double amin[3], amax[3], cell; i64 N[3]; i64 ngp;
  double *pc[3];   /* point coordinate per axis index:  amin + cell*i          (Grid.jl:87)  */
  i64 *cellof[3];  /* cell of the point per axis:       floor(N*(x-amin)/(amax-amin)) (Grid.jl:58) */
  i64 *cstart[3];  /* per axis: for cell c, points with cellof == c are [cstart[c], cstart[c+1])  (cellof is monotone) */
} ogrid;
static void grid_init(ogrid *g, const double amin[3], const double amax[3], const i64 N[3], double cell) {
  g->cell = cell; g->ngp = 1;
  for (int d = 0; d < 3; d++) {
    g->amin[d] = amin[d]; g->amax[d] = amax[d]; g->N[d] = N[d]; g->ngp *= (N[d] + 1);
    g->pc[d] = (double *)malloc(sizeof(double) * (size_t)(N[d] + 1));
    g->cellof[d] = (i64 *)malloc(sizeof(i64) * (size_t)(N[d] + 1));
    g->cstart[d] = (i64 *)malloc(sizeof(i64) * (size_t)(N[d] + 3));
    for (i64 i = 0; i <= N[d]; i++) {
      double x = amin[d] + cell * (double)i;
      g->pc[d][i] = x;
      g->cellof[d][i] = (i64)floor((double)N[d] * (x - amin[d]) / (amax[d] - amin[d]));
    }
    i64 p = 0;
    for (i64 c = 0; c <= N[d] + 1; c++) { while (p <= N[d] && g->cellof[d][p] < c) p++; g->cstart[d][c] = p; }
  }
}
static void grid_free(ogrid *g) { for (int d = 0; d < 3; d++) { free(g->pc[d]); free(g->cellof[d]); free(g->cstart[d]); } }
/* calculateMiniAABB_grid (Grid.jl:122-154): cell range of [lo-delta, hi+delta]; returns 0 if empty */
static int cell_range(const ogrid *g, const double lo[3], const double hi[3], double delta, i64 Imin[3], i64 Imax[3]) {
  for (int d = 0; d < 3; d++) {
    double a = lo[d] - delta, b = hi[d] + delta;
    double fmin = floor((double)g->N[d] * (a - g->amin[d]) / (g->amax[d] - g->amin[d]));
    double fmax = floor((double)g->N[d] * (b - g->amin[d]) / (g->amax[d] - g->amin[d]));
    if (fmin < 0) fmin = 0;
    if (fmax >= (double)g->N[d]) fmax = (double)g->N[d];
    if (fmin > fmax) return 0;
    Imin[d] = (i64)fmin; Imax[d] = (i64)fmax;
  }
  return 1;
}

/* ------------------------------------------------------------------------------------------------
 * Inverse isoparametric map  (restates FindLocalCoordinates.jl:16-107 as a converged Newton iteration;
 * decision-critical: fixed operation order, mirrored bit-for-bit by the CUDA device function)
 * returns 1 and xi on success, 0 and xi=(10,10,10) on failure (FindLocalCoordinates.jl:106)
 * ---------------------------------------------------------------------------------------------- */
/* monomial coefficients of a trilinear field v = A0 + A1 x + A2 e + A3 z + A4 xe + A5 ez + A6 zx + A7 xez from its nodal
 * values (node order of hex8_shape.jl:27-34); fixed operation order */
static void mono8(const double v[8], double A[8]) {
  double s01 = v[0] + v[1], d01 = v[1] - v[0], s32 = v[3] + v[2], d32 = v[2] - v[3];
  double s45 = v[4] + v[5], d45 = v[5] - v[4], s76 = v[7] + v[6], d76 = v[6] - v[7];
  double b0 = s01 + s32, b1 = d01 + d32, b2 = s32 - s01, b3 = d32 - d01;
  double t0 = s45 + s76, t1 = d45 + d76, t2 = s76 - s45, t3 = d76 - d45;
  A[0] = 0.125 * (b0 + t0); A[1] = 0.125 * (b1 + t1); A[2] = 0.125 * (b2 + t2); A[4] = 0.125 * (b3 + t3);
  A[3] = 0.125 * (t0 - b0); A[6] = 0.125 * (t1 - b1); A[5] = 0.125 * (t2 - b2); A[7] = 0.125 * (t3 - b3);
}
static int inverse_map_hex8(const double Xe[3][8], const double x[3], double xi[3]) {
  double A[3][8]; int affine = 1;
  for (int d = 0; d < 3; d++) { mono8(Xe[d], A[d]); if (A[d][4] != 0.0 || A[d][5] != 0.0 || A[d][6] != 0.0 || A[d][7] != 0.0) affine = 0; }
  xi[0] = xi[1] = xi[2] = 0.0;
  for (int it = 0; it < 50; it++) {
    double X = xi[0], E = xi[1], Z = xi[2], xe = X * E, ez = E * Z, zx = Z * X, xez = xe * Z;
    double r[3], J[3][3];
    for (int d = 0; d < 3; d++) {
      const double *a = A[d];
      double val = ((((((a[0] + a[1] * X) + a[2] * E) + a[3] * Z) + a[4] * xe) + a[5] * ez) + a[6] * zx) + a[7] * xez;
      r[d] = val - x[d];
      J[d][0] = ((a[1] + a[4] * E) + a[6] * Z) + a[7] * ez;
      J[d][1] = ((a[2] + a[4] * X) + a[5] * Z) + a[7] * zx;
      J[d][2] = ((a[3] + a[5] * E) + a[6] * X) + a[7] * xe;
    }
    /* solve J dx = r by the adjugate (cofactor) formula */
    double c00 = J[1][1] * J[2][2] - J[1][2] * J[2][1], c01 = J[1][2] * J[2][0] - J[1][0] * J[2][2], c02 = J[1][0] * J[2][1] - J[1][1] * J[2][0];
    double det = (J[0][0] * c00 + J[0][1] * c01) + J[0][2] * c02;
    if (!(fabs(det) > 0.0)) break;
    double c10 = J[0][2] * J[2][1] - J[0][1] * J[2][2], c11 = J[0][0] * J[2][2] - J[0][2] * J[2][0], c12 = J[0][1] * J[2][0] - J[0][0] * J[2][1];
    double c20 = J[0][1] * J[1][2] - J[0][2] * J[1][1], c21 = J[0][2] * J[1][0] - J[0][0] * J[1][2], c22 = J[0][0] * J[1][1] - J[0][1] * J[1][0];
    double d0 = ((c00 * r[0] + c10 * r[1]) + c20 * r[2]) / det;
    double d1 = ((c01 * r[0] + c11 * r[1]) + c21 * r[2]) / det;
    double d2 = ((c02 * r[0] + c12 * r[1]) + c22 * r[2]) / det;
    xi[0] = xi[0] - d0; xi[1] = xi[1] - d1; xi[2] = xi[2] - d2;
    double m = fmax(fabs(d0), fmax(fabs(d1), fabs(d2)));
    if (!(m < 1.0e3) || !(fmax(fabs(xi[0]), fmax(fabs(xi[1]), fabs(xi[2]))) < 1.0e3)) break;   /* diverged / NaN */
    if (affine) return 1;                    /* parallelepiped: the map is affine, one Newton step is exact */
    if (m < 1.0e-13) return 1;
  }
  xi[0] = xi[1] = xi[2] = 10.0; return 0;
}
/* TET4: FindLocalCoordinates.jl:110-149 -- lambda_234 = A \ b by the adjugate, coords = (l1, l2, l3) */
static int inverse_map_tet4(const double Xe[3][4], const double x[3], double lc[3]) {
  double A[3][3], b[3];
  for (int d = 0; d < 3; d++) { A[d][0] = Xe[d][1] - Xe[d][0]; A[d][1] = Xe[d][2] - Xe[d][0]; A[d][2] = Xe[d][3] - Xe[d][0]; b[d] = x[d] - Xe[d][0]; }
  double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  double det = (A[0][0] * c00 + A[0][1] * c01) + A[0][2] * c02;
  if (!(fabs(det) > 0.0)) { lc[0] = lc[1] = lc[2] = 10.0; return 0; }
  double c10 = A[0][2] * A[2][1] - A[0][1] * A[2][2], c11 = A[0][0] * A[2][2] - A[0][2] * A[2][0], c12 = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  double c20 = A[0][1] * A[1][2] - A[0][2] * A[1][1], c21 = A[0][2] * A[1][0] - A[0][0] * A[1][2], c22 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  double l2 = ((c00 * b[0] + c10 * b[1]) + c20 * b[2]) / det;
  double l3 = ((c01 * b[0] + c11 * b[1]) + c21 * b[2]) / det;
  double l4 = ((c02 * b[0] + c12 * b[1]) + c22 * b[2]) / det;
  double l1 = 1.0 - ((l2 + l3) + l4);
  /* validate_local_coords(TET4,[l1,l2,l3,l4]) (ElementTypes.jl:104-106): all >= 0 and sum <= 1.0 */
  if (l1 >= 0.0 && l2 >= 0.0 && l3 >= 0.0 && l4 >= 0.0 && (((l1 + l2) + l3) + l4) <= 1.0) { lc[0] = l1; lc[1] = l2; lc[2] = l3; return 1; }
  lc[0] = lc[1] = lc[2] = 10.0; return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Closest point on the in-element iso-surface, HEX8
 *   min ||x - Xe N(xi)||^2   s.t.  rho_e . N(xi) = rho_t,  -1 <= xi <= 1        (ComputeCoordsOnIso.jl:16-87)
 * The reference hands this to NLopt :LD_SLSQP from xi = 0.  Restated as a feasible-path SQP:
 *   phase 1  Newton-project xi=0 onto {g=0} inside the box (fallback: nearest iso-crossing of an element edge)
 *   phase 2  Newton steps in the tangent space of g on the current face of the box (exact Hessian of the
 *            Lagrangian, Gauss-Newton fallback), ratio test against the bounds, restoration onto g=0,
 *            Armijo on f; bounds are released by multiplier sign once the face problem has converged.
 * Converged to |step|_inf <= 1e-11 in xi.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { const double (*Xe)[8]; const double *re; const double *x; double rho_t; } isoprob;
typedef struct { double f, g, F[3], c[3], a[3], Hf[3][3], Hgn[3][3], Hg[3][3]; } isoeval;

static void iso_fg(const isoprob *p, const double xi[3], double *f, double *g, double a[3]) {
  double N[8], dN[8][3]; hex8_shape_d(xi, N, dN);
  double F[3];
  for (int d = 0; d < 3; d++) { double s = 0; for (int k = 0; k < 8; k++) s += p->Xe[d][k] * N[k]; F[d] = s - p->x[d]; }
  if (f) *f = F[0] * F[0] + F[1] * F[1] + F[2] * F[2];
  double s = 0; for (int k = 0; k < 8; k++) s += p->re[k] * N[k]; *g = s - p->rho_t;
  if (a) for (int j = 0; j < 3; j++) { double t = 0; for (int k = 0; k < 8; k++) t += p->re[k] * dN[k][j]; a[j] = t; }
}
static void iso_full(const isoprob *p, const double xi[3], isoeval *E) {
  double N[8], dN[8][3]; hex8_shape_d(xi, N, dN);
  double J[3][3];
  for (int d = 0; d < 3; d++) {
    double s = 0; for (int k = 0; k < 8; k++) s += p->Xe[d][k] * N[k]; E->F[d] = s - p->x[d];
    for (int j = 0; j < 3; j++) { double t = 0; for (int k = 0; k < 8; k++) t += p->Xe[d][k] * dN[k][j]; J[d][j] = t; }
  }
  E->f = E->F[0] * E->F[0] + E->F[1] * E->F[1] + E->F[2] * E->F[2];
  double s = 0; for (int k = 0; k < 8; k++) s += p->re[k] * N[k]; E->g = s - p->rho_t;
  for (int j = 0; j < 3; j++) {
    double t = 0; for (int k = 0; k < 8; k++) t += p->re[k] * dN[k][j]; E->a[j] = t;
    E->c[j] = 2.0 * (J[0][j] * E->F[0] + J[1][j] * E->F[1] + J[2][j] * E->F[2]);
  }
  /* mixed second derivatives of the trilinear basis: d2N_k/dxi_i dxi_j = 1/8 s_ki s_kj (1 + s_kl xi_l), l != i,j */
  double d2X[3][3] = {{0}}, d2g[3] = {0};      /* index m: 0 -> (xi,eta), 1 -> (eta,zeta), 2 -> (zeta,xi) */
  for (int k = 0; k < 8; k++) {
    const double *sg = HEX_SG[k];
    double h01 = 0.125 * sg[0] * sg[1] * (1 + sg[2] * xi[2]);
    double h12 = 0.125 * sg[1] * sg[2] * (1 + sg[0] * xi[0]);
    double h20 = 0.125 * sg[2] * sg[0] * (1 + sg[1] * xi[1]);
    for (int d = 0; d < 3; d++) { d2X[d][0] += p->Xe[d][k] * h01; d2X[d][1] += p->Xe[d][k] * h12; d2X[d][2] += p->Xe[d][k] * h20; }
    d2g[0] += p->re[k] * h01; d2g[1] += p->re[k] * h12; d2g[2] += p->re[k] * h20;
  }
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    E->Hgn[i][j] = 2.0 * (J[0][i] * J[0][j] + J[1][i] * J[1][j] + J[2][i] * J[2][j]);
    E->Hg[i][j] = 0.0; E->Hf[i][j] = E->Hgn[i][j];
  }
  double m01 = 2.0 * (E->F[0] * d2X[0][0] + E->F[1] * d2X[1][0] + E->F[2] * d2X[2][0]);
  double m12 = 2.0 * (E->F[0] * d2X[0][1] + E->F[1] * d2X[1][1] + E->F[2] * d2X[2][1]);
  double m20 = 2.0 * (E->F[0] * d2X[0][2] + E->F[1] * d2X[1][2] + E->F[2] * d2X[2][2]);
  E->Hf[0][1] += m01; E->Hf[1][0] += m01; E->Hf[1][2] += m12; E->Hf[2][1] += m12; E->Hf[2][0] += m20; E->Hf[0][2] += m20;
  E->Hg[0][1] = E->Hg[1][0] = d2g[0]; E->Hg[1][2] = E->Hg[2][1] = d2g[1]; E->Hg[2][0] = E->Hg[0][2] = d2g[2];
}
/* Newton restoration onto g = 0 moving only the variables with fix[i]==0; variables leaving the box are
 * clamped and fixed.  Returns 1 when |g| <= tol. */
static int iso_restore(const isoprob *p, double xi[3], int fix[3], double tolg) {
  for (int it = 0; it < 40; it++) {
    double g, a[3]; iso_fg(p, xi, NULL, &g, a);
    if (fabs(g) <= tolg) return 1;
    double den = 0; for (int i = 0; i < 3; i++) if (!fix[i]) den += a[i] * a[i];
    if (!(den > 0.0)) return 0;
    for (int i = 0; i < 3; i++) if (!fix[i]) {
      xi[i] -= g * a[i] / den;
      if (xi[i] >= 1.0) { xi[i] = 1.0; fix[i] = 1; } else if (xi[i] <= -1.0) { xi[i] = -1.0; fix[i] = -1; }
    }
  }
  return 0;
}
/* tangent Newton step on the free variables: min 1/2 d'Hd + c'd s.t. a_F'd = 0, d_fixed = 0 */
static void iso_tangent_step(const isoeval *E, const int fix[3], double lam, double d[3]) {
  int fr[3], nf = 0; for (int i = 0; i < 3; i++) if (!fix[i]) fr[nf++] = i;
  d[0] = d[1] = d[2] = 0.0;
  if (nf < 2) return;
  double H[3][3];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) H[i][j] = E->Hf[i][j] + lam * E->Hg[i][j];
  if (nf == 2) {
    int i = fr[0], j = fr[1]; double z[3] = {0, 0, 0}; z[i] = -E->a[j]; z[j] = E->a[i];
    double zz = z[i] * z[i] + z[j] * z[j]; if (!(zz > 0.0)) return;
    double kap = z[i] * (H[i][i] * z[i] + H[i][j] * z[j]) + z[j] * (H[j][i] * z[i] + H[j][j] * z[j]);
    double kgn = z[i] * (E->Hgn[i][i] * z[i] + E->Hgn[i][j] * z[j]) + z[j] * (E->Hgn[j][i] * z[i] + E->Hgn[j][j] * z[j]);
    if (!(kap > 1e-8 * kgn)) kap = kgn;
    if (!(kap > 0.0)) return;
    double t = -(z[i] * E->c[i] + z[j] * E->c[j]) / kap; d[i] = t * z[i]; d[j] = t * z[j]; return;
  }
  /* nf == 3: basis of null(a) built around the largest |a_k| */
  int k = 0; if (fabs(E->a[1]) > fabs(E->a[k])) k = 1; if (fabs(E->a[2]) > fabs(E->a[k])) k = 2;
  if (!(fabs(E->a[k]) > 0.0)) return;
  int u = (k + 1) % 3, v = (k + 2) % 3; double z1[3] = {0, 0, 0}, z2[3] = {0, 0, 0};
  z1[u] = E->a[k]; z1[k] = -E->a[u]; z2[v] = E->a[k]; z2[k] = -E->a[v];
  double Hz1[3], Hz2[3], Gz1[3], Gz2[3];
  for (int i = 0; i < 3; i++) {
    Hz1[i] = H[i][0] * z1[0] + H[i][1] * z1[1] + H[i][2] * z1[2]; Hz2[i] = H[i][0] * z2[0] + H[i][1] * z2[1] + H[i][2] * z2[2];
    Gz1[i] = E->Hgn[i][0] * z1[0] + E->Hgn[i][1] * z1[1] + E->Hgn[i][2] * z1[2]; Gz2[i] = E->Hgn[i][0] * z2[0] + E->Hgn[i][1] * z2[1] + E->Hgn[i][2] * z2[2];
  }
  double m11 = z1[0] * Hz1[0] + z1[1] * Hz1[1] + z1[2] * Hz1[2], m12 = z1[0] * Hz2[0] + z1[1] * Hz2[1] + z1[2] * Hz2[2], m22 = z2[0] * Hz2[0] + z2[1] * Hz2[1] + z2[2] * Hz2[2];
  double g11 = z1[0] * Gz1[0] + z1[1] * Gz1[1] + z1[2] * Gz1[2], g12 = z1[0] * Gz2[0] + z1[1] * Gz2[1] + z1[2] * Gz2[2], g22 = z2[0] * Gz2[0] + z2[1] * Gz2[1] + z2[2] * Gz2[2];
  double r1 = -(z1[0] * E->c[0] + z1[1] * E->c[1] + z1[2] * E->c[2]), r2 = -(z2[0] * E->c[0] + z2[1] * E->c[1] + z2[2] * E->c[2]);
  double det = m11 * m22 - m12 * m12, detg = g11 * g22 - g12 * g12;
  if (!(m11 > 1e-8 * g11 && det > 1e-8 * detg)) { m11 = g11; m12 = g12; m22 = g22; det = detg; }
  if (!(det > 0.0 && m11 > 0.0)) return;
  double y1 = (m22 * r1 - m12 * r2) / det, y2 = (m11 * r2 - m12 * r1) / det;
  for (int i = 0; i < 3; i++) d[i] = y1 * z1[i] + y2 * z2[i];
}
static int r2so_debug = 0;
API void r2so_set_debug(int v) { r2so_debug = v; }
static int project_iso_hex8(const double x[3], double rho_t, const double Xe[3][8], const double re[8], double xi[3], int *niter) {
  isoprob P = {Xe, re, x, rho_t};
  double gs = fabs(rho_t); for (int k = 0; k < 8; k++) if (fabs(re[k]) > gs) gs = fabs(re[k]); if (gs < 1.0) gs = 1.0;
  const double tolg = 1e-14 * gs, tolx = 1e-11, atol2 = 1e-24 * gs * gs;
  int fix[3] = {0, 0, 0};
  /* ---- phase 1: feasible point from xi = 0 ---- */
  xi[0] = xi[1] = xi[2] = 0.0;
  int ok = iso_restore(&P, xi, fix, tolg);
  if (!ok) {
    /* fallback: iso-crossing on an element edge closest to x (exists for every crossing element) */
    double best = INFINITY;
    for (int e = 0; e < 12; e++) {
      int a = HEX_EDGES[e][0], b = HEX_EDGES[e][1]; double ra = re[a] - rho_t, rb = re[b] - rho_t;
      if ((ra <= 0 && rb >= 0) || (ra >= 0 && rb <= 0)) {
        double t = (ra == rb) ? 0.5 : ra / (ra - rb), cand[3], q[3], N[8];
        for (int d = 0; d < 3; d++) cand[d] = HEX_SG[a][d] + t * (HEX_SG[b][d] - HEX_SG[a][d]);
        hex8_shape(cand, N);
        for (int d = 0; d < 3; d++) { double s = 0; for (int k = 0; k < 8; k++) s += Xe[d][k] * N[k]; q[d] = s - x[d]; }
        double dd = q[0] * q[0] + q[1] * q[1] + q[2] * q[2];
        if (dd < best) { best = dd; xi[0] = cand[0]; xi[1] = cand[1]; xi[2] = cand[2]; }
      }
    }
    if (!(best < INFINITY)) { xi[0] = xi[1] = xi[2] = 0.0; if (niter) *niter = -1; return 0; }
  }
  /* ---- phase 2 ---- */
  double lam = 0.0, dm = 0.0; int it, status = 0, stall = 0, force = 0;
  for (it = 0; it < 100 && status == 0; it++) {
    int bnd[3], tried[3] = {0, 0, 0};
    for (int i = 0; i < 3; i++) { bnd[i] = xi[i] >= 1.0 ? 1 : (xi[i] <= -1.0 ? -1 : 0); fix[i] = bnd[i]; }
    isoeval E; iso_full(&P, xi, &E);
    if (r2so_debug) printf("O it=%d xi=(%.15g %.15g %.15g) f=%.15g g=%.3e fix=(%d %d %d) force=%d\n", it, xi[0], xi[1], xi[2], E.f, E.g, fix[0], fix[1], fix[2], force);
    double d[3] = {0, 0, 0}; int have_step = 0;
    for (int pass = 0; pass < 8; pass++) {
      double num = 0, den = 0; for (int i = 0; i < 3; i++) if (!fix[i]) { num += E.a[i] * E.c[i]; den += E.a[i] * E.a[i]; }
      if (den > atol2) lam = -num / den;
      else {
        /* no usable free component of grad g: pick lambda inside the interval that makes every bound multiplier
         * valid, if there is one; otherwise the value that zeroes the multiplier of the largest |a_i| */
        double llo = -INFINITY, lhi = INFINITY; int kk = -1;
        for (int i = 0; i < 3; i++) if (fix[i]) {
          double as = E.a[i] * fix[i], cs = E.c[i] * fix[i];        /* need cs + lam*as <= 0 */
          if (as > 0) { double b = -cs / as; if (b < lhi) lhi = b; } else if (as < 0) { double b = -cs / as; if (b > llo) llo = b; }
          if (kk < 0 || fabs(E.a[i]) > fabs(E.a[kk])) kk = i;
        }
        if (llo <= lhi) lam = (llo > -INFINITY && lhi < INFINITY) ? 0.5 * (llo + lhi) : (llo > -INFINITY ? llo : (lhi < INFINITY ? lhi : 0.0));
        else if (kk >= 0 && E.a[kk] != 0.0) lam = -E.c[kk] / E.a[kk];
      }
      /* tangent (or, when grad g vanishes on the free set, unconstrained) Newton step */
      if (den > atol2) iso_tangent_step(&E, fix, lam, d);
      else {
        d[0] = d[1] = d[2] = 0.0;
        for (int i = 0; i < 3; i++) if (!fix[i]) {      /* diagonal Newton on the free variables (degenerate case) */
          double h = E.Hgn[i][i];                       /* Hess g has a zero diagonal */
          if (h > 0.0) d[i] = -E.c[i] / h;
        }
      }
      int refix = 0;
      for (int i = 0; i < 3; i++) if (!fix[i] && bnd[i] && d[i] * bnd[i] > 0.0) { if (fabs(d[i]) <= tolx) d[i] = 0.0; else { fix[i] = bnd[i]; refix = 1; } }      /* a numerically zero outward component is noise, not a reason to re-fix */
      if (refix) continue;
      dm = fmax(fabs(d[0]), fmax(fabs(d[1]), fabs(d[2])));
      if (r2so_debug) printf("O   pass=%d lam=%.15g d=(%.6e %.6e %.6e) fix=(%d %d %d) den=%.3e\n", pass, lam, d[0], d[1], d[2], fix[0], fix[1], fix[2], den);
      if (dm > tolx && !force) { have_step = 1; break; }
      /* converged on this face: release the bound with the most wrong-signed multiplier, if any */
      int worst = -1; double wv = 0.0;
      for (int i = 0; i < 3; i++) if (fix[i] && !tried[i]) {
        double gain = (E.c[i] + lam * E.a[i]) * (double)fix[i];   /* > 0: moving inward decreases the Lagrangian */
        if (gain > 1e-10 * (fabs(E.c[i]) + fabs(lam * E.a[i]) + 1e-300) && gain > wv) { wv = gain; worst = i; }
      }
      if (worst < 0) { status = 1; break; }
      fix[worst] = 0; tried[worst] = 1; force = 0;
    }
    if (status) break;
    if (!have_step) { status = 1; break; }
    /* ratio test against the box */
    double amax = 1.0; int blk = -1;
    for (int i = 0; i < 3; i++) if (!fix[i]) {
      /* ties between blocking bounds (within 1e-12) go to the lower index, so that round-off cannot choose the face */
      if (d[i] > 0 && xi[i] + d[i] > 1.0) { double t = (1.0 - xi[i]) / d[i]; if (t < amax * (1.0 - 1e-12)) { amax = t; blk = i; } }
      if (d[i] < 0 && xi[i] + d[i] < -1.0) { double t = (-1.0 - xi[i]) / d[i]; if (t < amax * (1.0 - 1e-12)) { amax = t; blk = i; } }
    }
    double slope = E.c[0] * d[0] + E.c[1] * d[1] + E.c[2] * d[2];
    if (!(slope < 0.0)) { force = 1; continue; }        /* no descent left on this face: go to the multiplier test */
    double alpha = amax; int acc = 0;
    for (int ls = 0; ls < 40; ls++) {
      double xt[3]; int fx[3] = {fix[0], fix[1], fix[2]};
      for (int i = 0; i < 3; i++) xt[i] = xi[i] + alpha * d[i];
      if (blk >= 0 && alpha == amax) { xt[blk] = d[blk] > 0 ? 1.0 : -1.0; fx[blk] = d[blk] > 0 ? 1 : -1; }
      for (int i = 0; i < 3; i++) { if (xt[i] >= 1.0) { xt[i] = 1.0; fx[i] = 1; } if (xt[i] <= -1.0) { xt[i] = -1.0; fx[i] = -1; } }
      if (iso_restore(&P, xt, fx, tolg)) {
        double ft, gt; iso_fg(&P, xt, &ft, &gt, NULL);
        /* steps below 1e-7 are in the quadratic-convergence regime of Newton: accept them without the Armijo test
         * (the decrease of f is below its evaluation noise there) */
        if (dm <= 1e-7 || ft <= E.f + 1e-4 * alpha * slope + 1e-15 * E.f) {
          if (E.f - ft <= 1e-15 * E.f) stall++; else stall = 0;
          xi[0] = xt[0]; xi[1] = xt[1]; xi[2] = xt[2]; acc = 1; break;
        }
      }
      alpha *= 0.5;
    }
    if (!acc) { if (dm < 1e-6) { force = 1; continue; } status = 2; break; }
    if (stall >= 3) { force = 1; stall = 0; }
  }
  if (niter) *niter = it;
  return status == 1;
}
static int project_iso_tet4(const double x[3], double rho_t, const double Xe[3][4], const double re[4], double xp[3]);
/* closest point on the iso-cut of one TET4 (test hook): Xe_colmajor = 3 x 4 column-major; returns 1 when a projection exists */
API int r2so_project_iso_tet4(const double *x, double rho_t, const double *Xe_colmajor, const double *re, double *xp) {
  double Xe[3][4]; for (int a = 0; a < 4; a++) for (int d = 0; d < 3; d++) Xe[d][a] = Xe_colmajor[3 * a + d];
  return project_iso_tet4(x, rho_t, (const double(*)[4])Xe, re, xp);
}
/* inverse isoparametric map of one HEX8 (test hook): Xe_colmajor = 3 x 8 column-major; returns 1 on success */
API int r2so_inverse_map_hex8(const double *x, const double *Xe_colmajor, double *xi) {
  double Xe[3][8]; for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) Xe[d][a] = Xe_colmajor[3 * a + d];
  return inverse_map_hex8((const double(*)[8])Xe, x, xi);
}
API int r2so_project_iso_hex8(const double *x, double rho_t, const double *Xe_colmajor, const double *re, double *xi, int *niter) {
  double Xe[3][8]; for (int a = 0; a < 8; a++) for (int d = 0; d < 3; d++) Xe[d][a] = Xe_colmajor[3 * a + d];
  return project_iso_hex8(x, rho_t, (const double(*)[8])Xe, re, xi, niter);
}

/* ------------------------------------------------------------------------------------------------
 * Closest point on the plane-cut of a TET4 (ComputeCoordsOnIso.jl:90-181)
 *   min ||x - Xe N(l)||^2 s.t. rho.N = rho_t, 0<=l_i<=1, sum l <= 1.  rho is affine on the tet, so the feasible set
 *   is the convex polygon {rho = rho_t} /\ tet and the (unique) minimiser is the projection of x on that polygon:
 *   the foot on the plane if it lies in the tet, else the closest point of the cut of the plane with a tet face.
 * ---------------------------------------------------------------------------------------------- */
static void closest_on_segment(const double a[3], const double b[3], const double x[3], double q[3]) {
  double e[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, ee = (e[0] * e[0] + e[1] * e[1]) + e[2] * e[2];
  double t = ee > 0 ? (((x[0] - a[0]) * e[0] + (x[1] - a[1]) * e[1]) + (x[2] - a[2]) * e[2]) / ee : 0.0;
  if (t < 0) t = 0;
  if (t > 1) t = 1;
  for (int d = 0; d < 3; d++) q[d] = a[d] + t * e[d];
}
static int tet_bary(const double Xe[3][4], const double p[3], double l[4]) {
  double A[3][3], b[3];
  for (int d = 0; d < 3; d++) { A[d][0] = Xe[d][1] - Xe[d][0]; A[d][1] = Xe[d][2] - Xe[d][0]; A[d][2] = Xe[d][3] - Xe[d][0]; b[d] = p[d] - Xe[d][0]; }
  double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  double det = (A[0][0] * c00 + A[0][1] * c01) + A[0][2] * c02;
  if (!(fabs(det) > 0.0)) return 0;
  double c10 = A[0][2] * A[2][1] - A[0][1] * A[2][2], c11 = A[0][0] * A[2][2] - A[0][2] * A[2][0], c12 = A[0][1] * A[2][0] - A[0][0] * A[2][1];
  double c20 = A[0][1] * A[1][2] - A[0][2] * A[1][1], c21 = A[0][2] * A[1][0] - A[0][0] * A[1][2], c22 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
  l[1] = ((c00 * b[0] + c10 * b[1]) + c20 * b[2]) / det;
  l[2] = ((c01 * b[0] + c11 * b[1]) + c21 * b[2]) / det;
  l[3] = ((c02 * b[0] + c12 * b[1]) + c22 * b[2]) / det;
  l[0] = 1.0 - ((l[1] + l[2]) + l[3]);
  return 1;
}
static int project_iso_tet4(const double x[3], double rho_t, const double Xe[3][4], const double re[4], double xp[3]) {
  /* physical gradient G of rho:  [x2-x1 x3-x1 x4-x1]^T G = [r2-r1 r3-r1 r4-r1] */
  double M[3][3], r[3] = {re[1] - re[0], re[2] - re[0], re[3] - re[0]}, G[3];
  for (int d = 0; d < 3; d++) { M[0][d] = Xe[d][1] - Xe[d][0]; M[1][d] = Xe[d][2] - Xe[d][0]; M[2][d] = Xe[d][3] - Xe[d][0]; }
  double det = det3((const double(*)[3])M);
  if (!(fabs(det) > 0.0)) return 0;
  for (int c = 0; c < 3; c++) { double B[3][3]; memcpy(B, M, sizeof(B)); for (int k = 0; k < 3; k++) B[k][c] = r[k]; G[c] = det3((const double(*)[3])B) / det; }
  double gg = (G[0] * G[0] + G[1] * G[1]) + G[2] * G[2];
  if (!(gg > 0.0)) return 0;
  double rx = re[0] + ((G[0] * (x[0] - Xe[0][0]) + G[1] * (x[1] - Xe[1][0])) + G[2] * (x[2] - Xe[2][0]));   /* rho extended to x */
  double s = (rx - rho_t) / gg, q[3] = {x[0] - s * G[0], x[1] - s * G[1], x[2] - s * G[2]}, l[4];
  if (tet_bary(Xe, q, l) && l[0] >= 0.0 && l[1] >= 0.0 && l[2] >= 0.0 && l[3] >= 0.0) { xp[0] = q[0]; xp[1] = q[1]; xp[2] = q[2]; return 1; }
  double best = INFINITY;
  for (int f = 0; f < 4; f++) {
    double P[3][3]; int np = 0;
    for (int e = 0; e < 3 && np < 3; e++) {
      int a = TET_ISN[f][e], b = TET_ISN[f][(e + 1) % 3]; double ra = re[a] - rho_t, rb = re[b] - rho_t;
      if ((ra <= 0 && rb > 0) || (ra > 0 && rb <= 0) || (ra < 0 && rb >= 0) || (ra >= 0 && rb < 0)) {
        double t = ra / (ra - rb);
        for (int d = 0; d < 3; d++) P[np][d] = Xe[d][a] + t * (Xe[d][b] - Xe[d][a]);
        np++;
      }
    }
    if (np < 2) continue;
    for (int i = 0; i + 1 < np; i++) {
      double c[3]; closest_on_segment(P[i], P[i + 1], x, c);
      double dd = ((x[0] - c[0]) * (x[0] - c[0]) + (x[1] - c[1]) * (x[1] - c[1])) + (x[2] - c[2]) * (x[2] - c[2]);
      if (dd < best) { best = dd; xp[0] = c[0]; xp[1] = c[1]; xp[2] = c[2]; }
    }
    if (np == 3) { double c[3]; closest_on_segment(P[2], P[0], x, c);
      double dd = ((x[0] - c[0]) * (x[0] - c[0]) + (x[1] - c[1]) * (x[1] - c[1])) + (x[2] - c[2]) * (x[2] - c[2]);
      if (dd < best) { best = dd; xp[0] = c[0]; xp[1] = c[1]; xp[2] = c[2]; } }
  }
  return best < INFINITY;
}

/* ------------------------------------------------------------------------------------------------
 * evalDistances (SignedDistances/sdfOnDensityField.jl:139-486) and helpers
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
  i64 nnp, nel; int nen, nes, nsn; const double *X; const i64 *IEN; ine_t ine;
} omesh;

static inline void write_value(double dist_tmp, const double xp[3], double *dist, double *xpo, i64 v) {
  /* WriteValue :44-57 / update_distance_parallel! :121-136 */
  if (fabs(dist_tmp) < fabs(dist[v])) { dist[v] = dist_tmp; if (xpo) { xpo[3 * v] = xp[0]; xpo[3 * v + 1] = xp[1]; xpo[3 * v + 2] = xp[2]; } }
}
/* is face sg of element el on the boundary?  (:511-519) intersection of the INE lists of its nodes has length 1 */
static int face_is_boundary(const omesh *m, i64 el, int sg) {
  const int *fn = m->nen == 8 ? HEX_ISN[sg] : TET_ISN[sg];
  i64 n0 = m->IEN[m->nen * el + fn[0]] - 1; int count = 0;
  for (i64 p = m->ine.ptr[n0]; p < m->ine.ptr[n0 + 1]; p++) {
    i64 e2 = m->ine.el[p]; int all = 1;
    for (int a = 1; a < m->nsn && all; a++) {
      i64 na = m->IEN[m->nen * el + fn[a]]; int found = 0;
      for (int b = 0; b < m->nen; b++) if (m->IEN[m->nen * e2 + b] == na) { found = 1; break; }
      all = found;
    }
    if (all) count++;
  }
  return count == 1;
}
/* barycentricCoordinates (TriangularMeshUtils.jl:1-24): 3x3 system, row of max |n| replaced by sum = 1,
 * solved by LU with partial pivoting (Julia's A \ b); fixed operation order (mirrored on the GPU) */
static void barycentric(const double x1[3], const double x2[3], const double x3[3], const double n[3], const double x[3], double lam[3]) {
  double A[3][3] = {
      {x1[1] * n[2] - x1[2] * n[1], x2[1] * n[2] - x2[2] * n[1], x3[1] * n[2] - x3[2] * n[1]},
      {x1[2] * n[0] - x1[0] * n[2], x2[2] * n[0] - x2[0] * n[2], x3[2] * n[0] - x3[0] * n[2]},
      {x1[0] * n[1] - x1[1] * n[0], x2[0] * n[1] - x2[1] * n[0], x3[0] * n[1] - x3[1] * n[0]}};
  double b[3] = {x[1] * n[2] - x[2] * n[1], x[2] * n[0] - x[0] * n[2], x[0] * n[1] - x[1] * n[0]};
  int im = 0; if (fabs(n[1]) > fabs(n[im])) im = 1; if (fabs(n[2]) > fabs(n[im])) im = 2;
  A[im][0] = A[im][1] = A[im][2] = 1.0; b[im] = 1.0;
  /* LU, partial pivoting */
  for (int k = 0; k < 2; k++) {
    int p = k; for (int r = k + 1; r < 3; r++) if (fabs(A[r][k]) > fabs(A[p][k])) p = r;
    if (p != k) { for (int c = 0; c < 3; c++) { double t = A[k][c]; A[k][c] = A[p][c]; A[p][c] = t; } double t = b[k]; b[k] = b[p]; b[p] = t; }
    for (int r = k + 1; r < 3; r++) {
      double l = A[r][k] / A[k][k];
      for (int c = k + 1; c < 3; c++) A[r][c] = A[r][c] - l * A[k][c];
      b[r] = b[r] - l * b[k];
    }
  }
  lam[2] = b[2] / A[2][2];
  lam[1] = (b[1] - A[1][2] * lam[2]) / A[1][1];
  lam[0] = ((b[0] - A[0][1] * lam[1]) - A[0][2] * lam[2]) / A[0][0];
}
/* IsProjectedOnFullSegment (:78-119) */
static int projected_on_full_segment(const omesh *m, i64 el, const double *rho_n, double rho_t, const double xp[3], const double x[3],
                                     double *dist, double *xpo, i64 v) {
  if (m->nen == 8) {
    double Xe[3][8], re[8], xi[3], N[8];
    for (int a = 0; a < 8; a++) { i64 n = m->IEN[8 * el + a] - 1; re[a] = rho_n[n]; for (int d = 0; d < 3; d++) Xe[d][a] = m->X[3 * n + d]; }
    inverse_map_hex8((const double(*)[8])Xe, xp, xi);
    if (!(fmax(fabs(xi[0]), fmax(fabs(xi[1]), fabs(xi[2]))) < 1.001)) return 0;
    hex8_shape(xi, N);
    double rho = N[0] * re[0]; for (int a = 1; a < 8; a++) rho = rho + N[a] * re[a];
    if (rho >= rho_t) { double dv[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]}; write_value(norm3(dv), xp, dist, xpo, v); return 1; }
    return 0;
  } else {
    double Xe[3][4], re[4], lc[3];
    for (int a = 0; a < 4; a++) { i64 n = m->IEN[4 * el + a] - 1; re[a] = rho_n[n]; for (int d = 0; d < 3; d++) Xe[d][a] = m->X[3 * n + d]; }
    if (!inverse_map_tet4((const double(*)[4])Xe, xp, lc)) return 0;   /* (10,10,10) fails validate_local_coords */
    /* validate_local_coords(TET4, local_coords[1:3]) && sum <= 1.001 (:97-99) */
    if (!(lc[0] >= 0 && lc[1] >= 0 && lc[2] >= 0 && ((lc[0] + lc[1]) + lc[2]) <= 1.0)) return 0;
    double l4 = 1.0 - ((lc[0] + lc[1]) + lc[2]);
    double rho = ((lc[0] * re[0] + lc[1] * re[1]) + lc[2] * re[2]) + l4 * re[3];
    if (rho >= rho_t) { double dv[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]}; write_value(norm3(dv), xp, dist, xpo, v); return 1; }
    return 0;
  }
}
/* process_triangle_projection! (:628-815) for one grid point */
static void triangle_point(const omesh *m, i64 el, const double *rho_n, double rho_t, int solid, const double Xt[3][3] /*[vertex][dim]*/,
                           const double Et[3][3], const double n[3], const double x[3], double *dist, double *xpo, i64 v) {
  double lam[3]; barycentric(Xt[0], Xt[1], Xt[2], n, x, lam);
  double xp[3]; int ok = 0;
  double lmin = lam[0]; if (lam[1] < lmin) lmin = lam[1]; if (lam[2] < lmin) lmin = lam[2];
  if (lmin >= 0.0) {
    for (int d = 0; d < 3; d++) xp[d] = (lam[0] * Xt[0][d] + lam[1] * Xt[1][d]) + lam[2] * Xt[2][d];
    double dv[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]}, dt = norm3(dv);
    if (solid) { if (fabs(dt) < fabs(dist[v])) { write_value(dt, xp, dist, xpo, v); ok = 1; } }
    else ok = projected_on_full_segment(m, el, rho_n, rho_t, xp, x, dist, xpo, v);
  } else {
    for (int j = 0; j < 3; j++) {
      double L = norm3(Et[j]);
      double u[3] = {Et[j][0] / L, Et[j][1] / L, Et[j][2] / L};
      double P = ((x[0] - Xt[j][0]) * u[0] + (x[1] - Xt[j][1]) * u[1]) + (x[2] - Xt[j][2]) * u[2];
      if (P >= 0 && P <= L) {
        for (int d = 0; d < 3; d++) xp[d] = Xt[j][d] + u[d] * P;
        double dv[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]}, dt = norm3(dv);
        if (solid) { if (fabs(dt) < fabs(dist[v])) { write_value(dt, xp, dist, xpo, v); ok = 1; } }
        else ok = projected_on_full_segment(m, el, rho_n, rho_t, xp, x, dist, xpo, v);
        if (ok) break;
      }
    }
  }
  if (!ok) {
    double dd[3];
    for (int j = 0; j < 3; j++) { double dv[3] = {x[0] - Xt[j][0], x[1] - Xt[j][1], x[2] - Xt[j][2]}; dd[j] = norm3(dv); }
    int idx = 0; if (dd[1] < dd[idx]) idx = 1; if (dd[2] < dd[idx]) idx = 2;
    for (int d = 0; d < 3; d++) xp[d] = Xt[idx][d];
    if (solid) write_value(dd[idx], xp, dist, xpo, v);
    else projected_on_full_segment(m, el, rho_n, rho_t, xp, x, dist, xpo, v);
  }
}
static void process_boundary_faces(const omesh *m, const ogrid *g, i64 el, const double *rho_n, double rho_t, double delta, int solid, double *dist, double *xpo) {
  for (int sg = 0; sg < m->nes; sg++) {
    if (!face_is_boundary(m, el, sg)) continue;
    const int *fn = m->nen == 8 ? HEX_ISN[sg] : TET_ISN[sg];
    double Xs[4][3], Xc[3];
    for (int a = 0; a < m->nsn; a++) for (int d = 0; d < 3; d++) Xs[a][d] = m->X[3 * (m->IEN[m->nen * el + fn[a]] - 1) + d];
    for (int d = 0; d < 3; d++) { double s = Xs[0][d]; for (int a = 1; a < m->nsn; a++) s = s + Xs[a][d]; Xc[d] = s / (double)m->nsn; }  /* mean(Xs, dims=2) */
    for (int a = 0; a < m->nsn; a++) {
      double Xt[3][3], Et[3][3], n[3], lo[3], hi[3];
      int a2 = (a + 1) % m->nsn;
      for (int d = 0; d < 3; d++) { Xt[0][d] = Xs[a][d]; Xt[1][d] = Xs[a2][d]; Xt[2][d] = Xc[d]; }
      for (int d = 0; d < 3; d++) { Et[0][d] = Xt[1][d] - Xt[0][d]; Et[1][d] = Xt[2][d] - Xt[1][d]; Et[2][d] = Xt[0][d] - Xt[2][d]; }  /* TriangularMeshUtils.jl:27-35 */
      n[0] = Et[0][1] * Et[1][2] - Et[0][2] * Et[1][1]; n[1] = Et[0][2] * Et[1][0] - Et[0][0] * Et[1][2]; n[2] = Et[0][0] * Et[1][1] - Et[0][1] * Et[1][0];
      double nn = norm3(n); n[0] = n[0] / nn; n[1] = n[1] / nn; n[2] = n[2] / nn;
      for (int d = 0; d < 3; d++) { lo[d] = fmin(Xt[0][d], fmin(Xt[1][d], Xt[2][d])); hi[d] = fmax(Xt[0][d], fmax(Xt[1][d], Xt[2][d])); }
      i64 I0[3], I1[3]; if (!cell_range(g, lo, hi, delta, I0, I1)) continue;
      for (i64 k = g->cstart[2][I0[2]]; k < g->cstart[2][I1[2] + 1]; k++)
        for (i64 j = g->cstart[1][I0[1]]; j < g->cstart[1][I1[1] + 1]; j++)
          for (i64 i = g->cstart[0][I0[0]]; i < g->cstart[0][I1[0] + 1]; i++) {
            i64 v = (k * (g->N[1] + 1) + j) * (g->N[0] + 1) + i;
            double x[3] = {g->pc[0][i], g->pc[1][j], g->pc[2][k]};
            triangle_point(m, el, rho_n, rho_t, solid, (const double(*)[3])Xt, (const double(*)[3])Et, n, x, dist, xpo, v);
          }
    }
  }
}
static void process_isocontour_element(const omesh *m, const ogrid *g, i64 el, const double *rho_n, double rho_t, double delta, double *dist, double *xpo, i64 *stats) {
  process_boundary_faces(m, g, el, rho_n, rho_t, delta, 0, dist, xpo);     /* :584 */
  double Xe8[3][8], Xe4[3][4], re[8], lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
  for (int a = 0; a < m->nen; a++) {
    i64 n = m->IEN[m->nen * el + a] - 1; re[a] = rho_n[n];
    for (int d = 0; d < 3; d++) { double c = m->X[3 * n + d]; if (m->nen == 8) Xe8[d][a] = c; else Xe4[d][a] = c; if (c < lo[d]) lo[d] = c; if (c > hi[d]) hi[d] = c; }
  }
  i64 I0[3], I1[3]; if (!cell_range(g, lo, hi, delta, I0, I1)) return;
  for (i64 k = g->cstart[2][I0[2]]; k < g->cstart[2][I1[2] + 1]; k++)
    for (i64 j = g->cstart[1][I0[1]]; j < g->cstart[1][I1[1] + 1]; j++)
      for (i64 i = g->cstart[0][I0[0]]; i < g->cstart[0][I1[0] + 1]; i++) {
        i64 v = (k * (g->N[1] + 1) + j) * (g->N[0] + 1) + i;
        double x[3] = {g->pc[0][i], g->pc[1][j], g->pc[2][k]}, xp[3];
        if (m->nen == 8) {
          double xi[3], N[8]; int nit = 0;
          int ok = project_iso_hex8(x, rho_t, (const double(*)[8])Xe8, re, xi, &nit);
          if (stats) { stats[0]++; stats[1] += nit > 0 ? nit : 0; if (!ok) stats[2]++; }
          hex8_shape(xi, N);
          for (int d = 0; d < 3; d++) { double s = 0; for (int a = 0; a < 8; a++) s += Xe8[d][a] * N[a]; xp[d] = s; }
        } else {
          int ok = project_iso_tet4(x, rho_t, (const double(*)[4])Xe4, re, xp);
          if (stats) { stats[0]++; if (!ok) stats[2]++; }
          if (!ok) continue;
        }
        double dv[3] = {x[0] - xp[0], x[1] - xp[1], x[2] - xp[2]};
        write_value(norm3(dv), xp, dist, xpo, v);
      }
}
/* nthreads == 1 : the reference's result under `julia -t 1` (single running-min buffer, elements ascending).
 * nthreads  > 1 : the reference's threaded scheme (:183-195,:457-461): one buffer per thread, contiguous element
 *                 chunks, merge by findmin -- used only for timing the CPU baseline. */
API int r2so_eval_distances(i64 nnp, const double *X, i64 nel, int nen, const i64 *IEN, const double *amin, const double *amax, const i64 *N, double cell,
                            const double *rho_n, double rho_t, double delta_factor, int nthreads, double *dist, double *xp, i64 *stats) {
  omesh m = {nnp, nel, nen, nen == 8 ? 6 : 4, nen == 8 ? 4 : 3, X, IEN, build_ine(nnp, nel, nen, IEN)};
  ogrid g; grid_init(&g, amin, amax, N, cell);
  double delta = delta_factor * cell;
  if (stats) stats[0] = stats[1] = stats[2] = 0;
  if (nthreads < 1) nthreads = 1;
  double **dl = (double **)malloc(sizeof(double *) * (size_t)nthreads), **xl = (double **)malloc(sizeof(double *) * (size_t)nthreads);
  for (int t = 0; t < nthreads; t++) {
    dl[t] = (double *)malloc(sizeof(double) * (size_t)g.ngp); for (i64 v = 0; v < g.ngp; v++) dl[t][v] = -BIG;
    xl[t] = xp ? (double *)calloc((size_t)g.ngp * 3, sizeof(double)) : NULL;
  }
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (int t = 0; t < nthreads; t++) {
    i64 e0 = nel * t / nthreads, e1 = nel * (t + 1) / nthreads;
    for (i64 el = e0; el < e1; el++) {
      double mn = DBL_MAX, mx = -DBL_MAX;
      for (int a = 0; a < nen; a++) { double r = rho_n[IEN[nen * el + a] - 1]; if (r < mn) mn = r; if (r > mx) mx = r; }
      if (mn >= rho_t) process_boundary_faces(&m, &g, el, rho_n, rho_t, delta, 1, dl[t], xl[t]);
      else if (mx > rho_t) process_isocontour_element(&m, &g, el, rho_n, rho_t, delta, dl[t], xl[t], nthreads == 1 ? stats : NULL);
    }
  }
  for (i64 v = 0; v < g.ngp; v++) {
    int bi = 0; double bd = fabs(dl[0][v]);
    for (int t = 1; t < nthreads; t++) if (fabs(dl[t][v]) < bd) { bd = fabs(dl[t][v]); bi = t; }
    dist[v] = bd;
    if (xp) { xp[3 * v] = xl[bi][3 * v]; xp[3 * v + 1] = xl[bi][3 * v + 1]; xp[3 * v + 2] = xl[bi][3 * v + 2]; }
  }
  for (int t = 0; t < nthreads; t++) { free(dl[t]); free(xl[t]); }
  free(dl); free(xl); grid_free(&g); free_ine(&m.ine); return 0;
}

/* ------------------------------------------------------------------------------------------------
 * Sign_Detection (SignedDistances/SignDetection.jl:6-81 HEX8, :88-268 TET4)
 * The HEX8 reference scans all nel AABBs per grid point; here candidates come from a per-cell element list
 * built in ascending element order, which yields the identical candidate sequence.
 * ---------------------------------------------------------------------------------------------- */
typedef struct { i64 *ptr; i64 *el; } bins_t;
API int r2so_sign_detection(i64 nnp, const double *X, i64 nel, int nen, const i64 *IEN, const double *amin, const double *amax, const i64 *N, double cell,
                            const double *rho_n, double rho_t, int nthreads, double *signs) {
  (void)nnp; ogrid g; grid_init(&g, amin, amax, N, cell);
  i64 nx = N[0] + 1, ny = N[1] + 1, nz = N[2] + 1;
  for (i64 v = 0; v < g.ngp; v++) signs[v] = -1.0;
  /* index ranges of grid points inside each element's closed AABB (HEX8: exact; TET4: the reference's padded cell range) */
  i64 *r0 = (i64 *)malloc(sizeof(i64) * 6 * (size_t)nel);
  i64 *cnt = (i64 *)calloc((size_t)g.ngp + 1, sizeof(i64));
  for (int pass = 0; pass < 2; pass++) {
    for (i64 e = 0; e < nel; e++) {
      i64 *r = r0 + 6 * e;
      if (pass == 0) {
        double lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int a = 0; a < nen; a++) for (int d = 0; d < 3; d++) { double c = X[3 * (IEN[nen * e + a] - 1) + d]; if (c < lo[d]) lo[d] = c; if (c > hi[d]) hi[d] = c; }
        for (int d = 0; d < 3; d++) {
          i64 n1 = N[d] + 1, a, b;
          if (nen == 8) {       /* all points with lo <= x <= hi (compute_aabb / is_point_inside_aabb, sdfOnDensityField.jl:60-69) */
            a = 0; while (a < n1 && g.pc[d][a] < lo[d]) a++;
            b = n1 - 1; while (b >= 0 && g.pc[d][b] > hi[d]) b--;
          } else {              /* create_grid_tetrahedra_mapping_TET4 :195-196 (1-based idx) and point_to_grid_index :256-268 */
            i64 mi = (i64)floor((lo[d] - amin[d]) / cell) - 1; if (mi < 1) mi = 1;
            i64 ma = (i64)ceil((hi[d] - amin[d]) / cell) + 1; if (ma > n1) ma = n1;
            /* a point with axis index p (0-based) looks up cell idx = clamp(floor((x-amin)/cell)+1, 1, n1) */
            a = n1; b = -1;
            for (i64 p = 0; p < n1; p++) { i64 idx = (i64)floor((g.pc[d][p] - amin[d]) / cell) + 1; if (idx < 1) idx = 1; if (idx > n1) idx = n1; if (idx >= mi && idx <= ma) { if (p < a) a = p; if (p > b) b = p; } }
          }
          r[2 * d] = a; r[2 * d + 1] = b;
        }
      }
      if (r[0] > r[1] || r[2] > r[3] || r[4] > r[5]) continue;
      for (i64 k = r[4]; k <= r[5]; k++) for (i64 j = r[2]; j <= r[3]; j++) for (i64 i = r[0]; i <= r[1]; i++) {
        i64 v = (k * ny + j) * nx + i;
        if (pass == 0) cnt[v + 1]++;
      }
    }
    if (pass == 0) for (i64 v = 0; v < g.ngp; v++) cnt[v + 1] += cnt[v];
  }
  i64 *lst = (i64 *)malloc(sizeof(i64) * (size_t)(cnt[g.ngp] > 0 ? cnt[g.ngp] : 1));
  i64 *fill = (i64 *)calloc((size_t)g.ngp, sizeof(i64));
  for (i64 e = 0; e < nel; e++) {
    i64 *r = r0 + 6 * e; if (r[0] > r[1] || r[2] > r[3] || r[4] > r[5]) continue;
    for (i64 k = r[4]; k <= r[5]; k++) for (i64 j = r[2]; j <= r[3]; j++) for (i64 i = r[0]; i <= r[1]; i++) { i64 v = (k * ny + j) * nx + i; lst[cnt[v] + fill[v]++] = e; }
  }
  free(fill);
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (i64 v = 0; v < g.ngp; v++) {
    i64 nc = cnt[v + 1] - cnt[v]; if (nc == 0) continue;
    i64 i = v % nx, j = (v / nx) % ny, k = v / (nx * ny);
    double x[3] = {g.pc[0][i], g.pc[1][j], g.pc[2][k]};
    if (nen == 8) {
      double mx = -DBL_MAX;
      for (i64 c = 0; c < nc; c++) { i64 e = lst[cnt[v] + c]; for (int a = 0; a < 8; a++) { double r = rho_n[IEN[8 * e + a] - 1]; if (r > mx) mx = r; } }
      if (mx < rho_t) continue;                                            /* :36 */
      double max_local = 10.0;
      for (i64 c = 0; c < nc; c++) {
        i64 e = lst[cnt[v] + c]; double Xe[3][8], re[8], xi[3], Nn[8];
        for (int a = 0; a < 8; a++) { i64 n = IEN[8 * e + a] - 1; re[a] = rho_n[n]; for (int d = 0; d < 3; d++) Xe[d][a] = X[3 * n + d]; }
        inverse_map_hex8((const double(*)[8])Xe, x, xi);
        double mn = fmax(fabs(xi[0]), fmax(fabs(xi[1]), fabs(xi[2])));
        if (mn < 1.01 && max_local > mn) {                                 /* :48 */
          hex8_shape(xi, Nn);
          double rho = Nn[0] * re[0]; for (int a = 1; a < 8; a++) rho = rho + Nn[a] * re[a];
          if (rho >= rho_t) signs[v] = 1.0;
          if (mn < 0.95) break;                                            /* :51-59 */
          max_local = mn;
        }
      }
    } else {
      for (i64 c = 0; c < nc; c++) {
        i64 e = lst[cnt[v] + c]; double Xe[3][4], re[4], lo[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, hi[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
        for (int a = 0; a < 4; a++) { i64 n = IEN[4 * e + a] - 1; re[a] = rho_n[n]; for (int d = 0; d < 3; d++) { double cc = X[3 * n + d]; Xe[d][a] = cc; if (cc < lo[d]) lo[d] = cc; if (cc > hi[d]) hi[d] = cc; } }
        /* is_point_in_tetrahedron :220-242 (tol 1e-10): AABB test, then 4x4 barycentric solve (same solution as the 3x3 reduced system) */
        const double tol = 1e-10; int out = 0;
        for (int d = 0; d < 3; d++) if (x[d] < lo[d] - tol || x[d] > hi[d] + tol) out = 1;
        if (out) continue;
        double A[3][3], b[3];
        for (int d = 0; d < 3; d++) { A[d][0] = Xe[d][1] - Xe[d][0]; A[d][1] = Xe[d][2] - Xe[d][0]; A[d][2] = Xe[d][3] - Xe[d][0]; b[d] = x[d] - Xe[d][0]; }
        double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2], c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
        double det = (A[0][0] * c00 + A[0][1] * c01) + A[0][2] * c02; if (!(fabs(det) > 0.0)) continue;
        double c10 = A[0][2] * A[2][1] - A[0][1] * A[2][2], c11 = A[0][0] * A[2][2] - A[0][2] * A[2][0], c12 = A[0][1] * A[2][0] - A[0][0] * A[2][1];
        double c20 = A[0][1] * A[1][2] - A[0][2] * A[1][1], c21 = A[0][2] * A[1][0] - A[0][0] * A[1][2], c22 = A[0][0] * A[1][1] - A[0][1] * A[1][0];
        double l2 = ((c00 * b[0] + c10 * b[1]) + c20 * b[2]) / det, l3 = ((c01 * b[0] + c11 * b[1]) + c21 * b[2]) / det, l4 = ((c02 * b[0] + c12 * b[1]) + c22 * b[2]) / det;
        double l1 = 1.0 - ((l2 + l3) + l4);
        if (!(l1 >= -tol && l2 >= -tol && l3 >= -tol && l4 >= -tol && l1 <= 1.0 + tol && l2 <= 1.0 + tol && l3 <= 1.0 + tol && l4 <= 1.0 + tol)) continue;
        double lc[3]; if (!inverse_map_tet4((const double(*)[4])Xe, x, lc)) continue;    /* found == false :132 */
        double l4b = 1.0 - ((lc[0] + lc[1]) + lc[2]);
        double rho = ((lc[0] * re[0] + lc[1] * re[1]) + lc[2] * re[2]) + l4b * re[3];
        if (rho >= rho_t) { signs[v] = 1.0; break; }
      }
    }
  }
  free(lst); free(cnt); free(r0); grid_free(&g); return 0;
}

/* ------------------------------------------------------------------------------------------------
 * remove_sdf_artifacts! (SignedDistances/SdfArtifactRemoval.jl:134-245): sequential union-find, 6-connectivity
 * ---------------------------------------------------------------------------------------------- */
static i64 uf_find(i64 *p, i64 x) { while (p[x] != x) { p[x] = p[p[x]]; x = p[x]; } return x; }
API int r2so_remove_artifacts(double *sdf, const i64 *N, double threshold, double min_ratio, i64 *flipped) {
  i64 nx = N[0] + 1, ny = N[1] + 1, nz = N[2] + 1, n = nx * ny * nz; *flipped = 0;
  i64 *par = (i64 *)malloc(sizeof(i64) * (size_t)n), *sz = (i64 *)calloc((size_t)n, sizeof(i64)); i64 interior = 0;
  for (i64 v = 0; v < n; v++) { par[v] = v; if (sdf[v] >= threshold) interior++; }
  if (interior == 0) { free(par); free(sz); return 0; }
  for (i64 k = 0; k < nz; k++) for (i64 j = 0; j < ny; j++) for (i64 i = 0; i < nx; i++) {
    i64 v = (k * ny + j) * nx + i; if (!(sdf[v] >= threshold)) continue;
    if (i + 1 < nx && sdf[v + 1] >= threshold) { i64 a = uf_find(par, v), b = uf_find(par, v + 1); if (a != b) par[a > b ? a : b] = a > b ? b : a; }
    if (j + 1 < ny && sdf[v + nx] >= threshold) { i64 a = uf_find(par, v), b = uf_find(par, v + nx); if (a != b) par[a > b ? a : b] = a > b ? b : a; }
    if (k + 1 < nz && sdf[v + nx * ny] >= threshold) { i64 a = uf_find(par, v), b = uf_find(par, v + nx * ny); if (a != b) par[a > b ? a : b] = a > b ? b : a; }
  }
  i64 largest = 0, lroot = -1;
  for (i64 v = 0; v < n; v++) if (sdf[v] >= threshold) { i64 r = uf_find(par, v); sz[r]++; }
  for (i64 v = 0; v < n; v++) if (sz[v] > largest) { largest = sz[v]; lroot = v; }
  /* min_component_size = max(1, round(Int, ratio*largest)) with round-half-to-even (:206) */
  double q = min_ratio * (double)largest; i64 ms = (i64)nearbyint(q); if (ms < 1) ms = 1;
  i64 nf = 0;
  for (i64 v = 0; v < n; v++) par[v] = (sdf[v] >= threshold) ? uf_find(par, v) : -1;      /* flatten first, then flip */
  for (i64 v = 0; v < n; v++) if (par[v] >= 0 && par[v] != lroot && sz[par[v]] < ms) { sdf[v] = -fabs(sdf[v]); nf++; }
  *flipped = nf; free(par); free(sz); return 0;
}

/* ------------------------------------------------------------------------------------------------
 * calculate_volume_from_sdf (SdfSmoothing/CalcVolumeFromSDF.jl:26-125), Float32; per-cell value exactly as the
 * reference, cell values summed in double (the reference's Atomic{Float32} order is nondeterministic)
 * ---------------------------------------------------------------------------------------------- */
static float cell_volume_f32(const float c[8], float iso, const float *gp, const float *w, int n, float elem_vol, float jac) {
  float mn = c[0], mx = c[0]; for (int a = 1; a < 8; a++) { if (c[a] < mn) mn = c[a]; if (c[a] > mx) mx = c[a]; }
  if (mx < iso) return 0.0f;
  if (mn >= iso) return elem_vol;
  float part = 0.0f;
  for (int kq = 0; kq < n; kq++) { float zeta = (gp[kq] + 1) / 2;
    for (int jq = 0; jq < n; jq++) { float eta = (gp[jq] + 1) / 2;
      for (int iq = 0; iq < n; iq++) { float xi = (gp[iq] + 1) / 2;
        float c00 = c[0] * (1.0f - xi) + c[1] * xi, c01 = c[4] * (1.0f - xi) + c[5] * xi;
        float c10 = c[2] * (1.0f - xi) + c[3] * xi, c11 = c[6] * (1.0f - xi) + c[7] * xi;
        float c0 = c00 * (1.0f - eta) + c10 * eta, c1 = c01 * (1.0f - eta) + c11 * eta;
        float ps = c0 * (1.0f - zeta) + c1 * zeta;
        if (ps >= iso) { float wt = w[iq] * w[jq] * w[kq]; part += wt * jac; }
      } } }
  return part;
}
/* c[] order: c000,c100,c010,c110,c001,c101,c011,c111 */
API double r2so_volume_from_sdf(const float *sdf, i64 nx, i64 ny, i64 nz, float edge, float iso, int order, int nthreads) {
  double gpd[32], wd[32]; float gp[32], w[32]; gauss_legendre(order, gpd, wd);
  for (int i = 0; i < order; i++) { gp[i] = (float)gpd[i]; w[i] = (float)wd[i]; }
  float ev = edge * edge * edge, jac = ev / 8.0f; double total = 0.0;
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static) reduction(+ : total)
#endif
  for (i64 k = 0; k < nz - 1; k++) {
    double plane = 0.0;
    for (i64 j = 0; j < ny - 1; j++) for (i64 i = 0; i < nx - 1; i++) {
      i64 b = (k * ny + j) * nx + i;
      float c[8] = {sdf[b], sdf[b + 1], sdf[b + nx], sdf[b + nx + 1], sdf[b + nx * ny], sdf[b + nx * ny + 1], sdf[b + nx * ny + nx], sdf[b + nx * ny + nx + 1]};
      plane += (double)cell_volume_f32(c, iso, gp, w, order, ev, jac);
    }
    total += plane;
  }
  return total;
}

/* ------------------------------------------------------------------------------------------------
 * RBFs_smoothing (SdfSmoothing/RBFs4Smoothing.jl:321-377)
 * ---------------------------------------------------------------------------------------------- */
static void f32_range(float a, float b, i64 n, float *out) {   /* range(Float32(a), Float32(b), length=n) (:41-43): Float64 intermediate */
  for (i64 i = 0; i < n; i++) out[i] = (n > 1) ? (float)((double)a + (double)i * (((double)b - (double)a) / (double)(n - 1))) : a;
  if (n > 1) out[n - 1] = b;
}
/* mode: 0 = faithful (kernel values from the Float32 coordinates, as the reference computes them)
 *       1 = ideal lattice (exp(-m), m = squared integer offset; what the CUDA stencil uses)                     */
typedef struct {
  i64 nx, ny, nz; int noff; int off[128][3]; double sigma; float maxd; int mode; float *cx, *cy, *cz;
} rbf_t;
static float rbf_kval(const rbf_t *R, i64 i, i64 j, i64 k, int o) {
  int di = R->off[o][0], dj = R->off[o][1], dk = R->off[o][2];
  if (R->mode == 1) { double m = (double)(di * di + dj * dj + dk * dk); return (float)exp(-m); }
  float dx = R->cx[i] - R->cx[i + di], dy = R->cy[j] - R->cy[j + dj], dz = R->cz[k] - R->cz[k + dk];
  float r = sqrtf(dx * dx + dy * dy + dz * dz);                     /* :104 Float32 */
  double q = (double)r / R->sigma; double val = exp(-(q * q));      /* :105 Float64 */
  return (float)val;
}
static void rbf_matvec(const rbf_t *R, const float *u, float *c, int nthreads) {
  /* mul!(c, K, u) for the CSC matrix of compute_sparse_kernel_matrix (:142-176): per row, Float32 accumulation in ascending column order */
  i64 nx = R->nx, ny = R->ny, nz = R->nz;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (i64 k = 0; k < nz; k++) for (i64 j = 0; j < ny; j++) for (i64 i = 0; i < nx; i++) {
    float acc = 0.0f;
    for (int o = 0; o < R->noff; o++) {      /* offsets are stored in ascending linear-index order */
      i64 ii = i + R->off[o][0], jj = j + R->off[o][1], kk = k + R->off[o][2];
      if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
      acc += rbf_kval(R, i, j, k, o) * u[(kk * ny + jj) * nx + ii];
    }
    c[(k * ny + j) * nx + i] = acc;
  }
}
static int cmp_off(const void *a, const void *b) {
  const int *p = (const int *)a, *q = (const int *)b;
  if (p[2] != q[2]) return p[2] - q[2]; if (p[1] != q[1]) return p[1] - q[1]; return p[0] - q[0];
}
/* outputs: fine_sdf (prod(N*smooth+1) floats, x fastest), th, final volume, cg iterations; weights_out/lsf_out optional (ngp floats) */
API int r2so_rbf_smoothing(const double *sdf, const double *amin, const double *amax, const i64 *N, double cell, int is_interp, int smooth, double rbf_cut,
                           double target_volume, int mode, int nthreads, float *fine_sdf, float *th_out, float *vol_out, int *cg_iters, float *weights_out, float *lsf_out) {
  i64 nx = N[0] + 1, ny = N[1] + 1, nz = N[2] + 1, n = nx * ny * nz;
  if (nthreads < 1) nthreads = 1;
  /* process_vector (:15-22) */
  float *s = (float *)malloc(sizeof(float) * (size_t)n); float maxv = -1.0f;
  for (i64 v = 0; v < n; v++) { s[v] = (float)sdf[v]; float a = fabsf(s[v]); if (a < 1.0e9f && a > maxv) maxv = a; }
  if (maxv < 0.0f) { free(s); return 2; }
  for (i64 v = 0; v < n; v++) {
    float a = fabsf(s[v]); float big = 1.0e10f, rt = sqrtf(FLT_EPSILON);
    if (fabsf(a - big) <= rt * fmaxf(a, big)) s[v] = (s[v] > 0 ? 1.0f : (s[v] < 0 ? -1.0f : 0.0f)) * maxv;   /* isapprox, default rtol */
  }
  rbf_t R; R.nx = nx; R.ny = ny; R.nz = nz; R.sigma = cell; R.mode = mode;
  R.maxd = (float)sqrt(-log(rbf_cut) * cell * cell);                                    /* :221 */
  R.cx = (float *)malloc(sizeof(float) * (size_t)nx); R.cy = (float *)malloc(sizeof(float) * (size_t)ny); R.cz = (float *)malloc(sizeof(float) * (size_t)nz);
  f32_range((float)amin[0], (float)amax[0], nx, R.cx); f32_range((float)amin[1], (float)amax[1], ny, R.cy); f32_range((float)amin[2], (float)amax[2], nz, R.cz);
  double rad = sqrt(-log(rbf_cut)); int ir = (int)floor(rad) + 1; R.noff = 0;
  for (int dk = -ir; dk <= ir; dk++) for (int dj = -ir; dj <= ir; dj++) for (int di = -ir; di <= ir; di++) {
    double m = di * di + dj * dj + dk * dk;
    if (exp(-m) > rbf_cut && R.noff < 128) { R.off[R.noff][0] = di; R.off[R.noff][1] = dj; R.off[R.noff][2] = dk; R.noff++; }
  }
  qsort(R.off, (size_t)R.noff, sizeof(R.off[0]), cmp_off);
  /* weights (:351-353): cg(K, raw) with IterativeSolvers defaults (x0 = 0, reltol = sqrt(eps(Float32)), maxiter = n) */
  float *wgt = (float *)malloc(sizeof(float) * (size_t)n); int iters = 0;
  if (is_interp) {
    float *r = (float *)malloc(sizeof(float) * (size_t)n), *u = (float *)calloc((size_t)n, sizeof(float)), *c = (float *)malloc(sizeof(float) * (size_t)n);
    double rr = 0; for (i64 v = 0; v < n; v++) { wgt[v] = 0.0f; r[v] = s[v]; rr += (double)r[v] * (double)r[v]; }
    float residual = (float)sqrt(rr), prev = 1.0f, tol = sqrtf(FLT_EPSILON) * residual;
    while (iters < n && !(residual <= tol)) {
      float beta = residual * residual / (prev * prev);
      for (i64 v = 0; v < n; v++) u[v] = r[v] + beta * u[v];
      rbf_matvec(&R, u, c, nthreads);
      double uc = 0; for (i64 v = 0; v < n; v++) uc += (double)u[v] * (double)c[v];
      float alpha = residual * residual / (float)uc;
      rr = 0; for (i64 v = 0; v < n; v++) { wgt[v] += alpha * u[v]; r[v] -= alpha * c[v]; rr += (double)r[v] * (double)r[v]; }
      prev = residual; residual = (float)sqrt(rr); iters++;
    }
    free(r); free(u); free(c);
  } else memcpy(wgt, s, sizeof(float) * (size_t)n);
  if (cg_iters) *cg_iters = iters;
  if (weights_out) memcpy(weights_out, wgt, sizeof(float) * (size_t)n);
  /* LSF on the coarse grid (:357) and on the fine grid (:363): sum over coarse nodes within max_distance,
   * ascending distance (ties: ascending index), result = Float32(result + w*exp(-(d/sigma)^2)) (:241) */
  float *lsf = (float *)malloc(sizeof(float) * (size_t)n);
  for (int lvl = 0; lvl < 2; lvl++) {
    int sm = lvl == 0 ? 1 : smooth; i64 fx = N[0] * sm + 1, fy = N[1] * sm + 1, fz = N[2] * sm + 1;
    float *out = lvl == 0 ? lsf : fine_sdf;
    float dxf = ((float)amax[0] - (float)amin[0]) / (float)(fx - 1);                      /* :65 */
    /* per phase (sub-cell position) tap list in half-lattice units, sorted by distance */
    int ntap[8]; int tap[8][160][3]; float tapw[8][160];
    for (int ph = 0; ph < sm * sm * sm; ph++) {
      int px = ph % sm, py = (ph / sm) % sm, pz = ph / (sm * sm), cntp = 0; double dd2[160];
      for (int dk = -4; dk <= 4; dk++) for (int dj = -4; dj <= 4; dj++) for (int di = -4; di <= 4; di++) {
        double ox = di - (double)px / sm, oy = dj - (double)py / sm, oz = dk - (double)pz / sm, m = ox * ox + oy * oy + oz * oz;
        float dist = (float)(sqrt(m) * (double)cell);
        if (dist <= R.maxd && cntp < 160) { tap[ph][cntp][0] = di; tap[ph][cntp][1] = dj; tap[ph][cntp][2] = dk; dd2[cntp] = m; tapw[ph][cntp] = (float)exp(-m); cntp++; }
      }
      for (int a = 1; a < cntp; a++) {  /* insertion sort by (distance, linear offset) */
        int t0 = tap[ph][a][0], t1 = tap[ph][a][1], t2 = tap[ph][a][2]; double dm = dd2[a]; float tw = tapw[ph][a]; int b = a - 1;
        while (b >= 0 && dd2[b] > dm) { tap[ph][b + 1][0] = tap[ph][b][0]; tap[ph][b + 1][1] = tap[ph][b][1]; tap[ph][b + 1][2] = tap[ph][b][2]; dd2[b + 1] = dd2[b]; tapw[ph][b + 1] = tapw[ph][b]; b--; }
        tap[ph][b + 1][0] = t0; tap[ph][b + 1][1] = t1; tap[ph][b + 1][2] = t2; dd2[b + 1] = dm; tapw[ph][b + 1] = tw;
      }
      ntap[ph] = cntp;
    }
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
    for (i64 k = 0; k < fz; k++) for (i64 j = 0; j < fy; j++) for (i64 i = 0; i < fx; i++) {
      int ph = (int)((i % sm) + sm * ((j % sm) + sm * (k % sm))); i64 ci = i / sm, cj = j / sm, ck = k / sm;
      float acc = 0.0f;
      float fxp = (float)amin[0] + (float)i * dxf, fyp = (float)amin[1] + (float)j * dxf, fzp = (float)amin[2] + (float)k * dxf;   /* :66-71 */
      int ncnt = 0;
      for (int t = 0; t < ntap[ph]; t++) {
        i64 ii = ci + tap[ph][t][0], jj = cj + tap[ph][t][1], kk = ck + tap[ph][t][2];
        if (ii < 0 || ii >= nx || jj < 0 || jj >= ny || kk < 0 || kk >= nz) continue;
        if (++ncnt > 124) break;                                                          /* knn k = 124 (:238) */
        double kv;
        if (mode == 1) kv = (double)tapw[ph][t];
        else {
          float ddx = fxp - R.cx[ii], ddy = fyp - R.cy[jj], ddz = fzp - R.cz[kk]; float dist = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz);
          if (lvl == 0) { ddx = R.cx[ci] - R.cx[ii]; ddy = R.cy[cj] - R.cy[jj]; ddz = R.cz[ck] - R.cz[kk]; dist = sqrtf(ddx * ddx + ddy * ddy + ddz * ddz); }
          if (!(dist <= R.maxd)) continue;
          double q = (double)dist / cell; kv = exp(-(q * q));
        }
        acc = (float)((double)acc + (double)wgt[(kk * ny + jj) * nx + ii] * kv);
      }
      out[(k * fy + j) * fx + i] = acc;
    }
  }
  if (lsf_out) memcpy(lsf_out, lsf, sizeof(float) * (size_t)n);
  /* LS_Threshold (:265-300) on the coarse LSF */
  float edge; { float ex = R.cx[1] - R.cx[0]; edge = sqrtf(ex * ex); }                   /* norm(grid[2,1,1]-grid[1,1,1]) */
  float lo = lsf[0], hi = lsf[0]; for (i64 v = 1; v < n; v++) { if (lsf[v] < lo) lo = lsf[v]; if (lsf[v] > hi) hi = lsf[v]; }
  float *sh = (float *)malloc(sizeof(float) * (size_t)n); double eps = 1.0; int nb = 0; float th = 0.0f;
  while (nb < 40 && eps > 1.0e-4) {
    th = (lo + hi) / 2;
    for (i64 v = 0; v < n; v++) sh[v] = lsf[v] - th;
    float cur = (float)r2so_volume_from_sdf(sh, nx, ny, nz, edge, 0.0f, 9, nthreads);
    eps = fabs(target_volume - (double)cur);
    if ((double)cur > target_volume) lo = th; else hi = th;
    nb++;
  }
  free(sh);
  float tho = -th; if (th_out) *th_out = tho;
  i64 fx = N[0] * smooth + 1, fy = N[1] * smooth + 1, fz = N[2] * smooth + 1, nf = fx * fy * fz;
  for (i64 v = 0; v < nf; v++) fine_sdf[v] = fine_sdf[v] + tho;                          /* :366 */
  if (vol_out) {
    float dxf = ((float)amax[0] - (float)amin[0]) / (float)(fx - 1); float x0 = (float)amin[0], x1 = (float)amin[0] + 1.0f * dxf; float e = sqrtf((x1 - x0) * (x1 - x0));
    *vol_out = (float)r2so_volume_from_sdf(fine_sdf, fx, fy, fz, e, 0.0f, 9, nthreads);
  }
  free(lsf); free(wgt); free(s); free(R.cx); free(R.cy); free(R.cz); return 0;
}

API int r2so_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
