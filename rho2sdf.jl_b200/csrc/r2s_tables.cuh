// r2s_tables.cuh -- element topology (ElementTypes/ElementTypes.jl:15-78, 0-based) and Gauss-Legendre tables
#pragma once
#include <math.h>

// face -> local nodes (ISN)
static __constant__ int c_hex_isn[6][4] = {{0, 3, 2, 1}, {0, 1, 5, 4}, {1, 2, 6, 5}, {2, 3, 7, 6}, {3, 0, 4, 7}, {4, 5, 6, 7}};
static __constant__ int c_tet_isn[4][3] = {{0, 2, 1}, {0, 1, 3}, {1, 2, 3}, {0, 3, 2}};
// natural coordinates of the HEX8 nodes (hex8_shape.jl:27-34)
static __constant__ double c_hex_sg[8][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1}};
static __constant__ int c_hex_edges[12][2] = {{0, 1}, {1, 2}, {2, 3}, {3, 0}, {4, 5}, {5, 6}, {6, 7}, {7, 4}, {0, 4}, {1, 5}, {2, 6}, {3, 7}};

// Gauss-Legendre rule of order n (what FastGaussQuadrature.gausslegendre(n) returns): Newton on P_n in long double,
// nodes ascending, symmetrised.  Host side; the tables travel to kernels by value.
struct GaussTab { double x[32]; double w[32]; int n; };      // orders up to 32 (the reference uses 3, 9, 15 and, in its convergence tests, 20)
static inline GaussTab gauss_legendre_host(int n) {
  GaussTab t; t.n = n;
  for (int i = 0; i < n; i++) {
    long double z = cosl(3.14159265358979323846264338327950288L * (i + 0.75L) / (n + 0.5L)), pp = 0;
    for (int it = 0; it < 100; it++) {
      long double p1 = 1, p2 = 0;
      for (int j = 0; j < n; j++) { long double p3 = p2; p2 = p1; p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1); }
      pp = n * (z * p1 - p2) / (z * z - 1);
      long double dz = p1 / pp; z -= dz;
      if (fabsl(dz) < 1e-19L) break;
    }
    long double p1 = 1, p2 = 0;
    for (int j = 0; j < n; j++) { long double p3 = p2; p2 = p1; p1 = ((2 * j + 1) * z * p2 - j * p3) / (j + 1); }
    pp = n * (z * p1 - p2) / (z * z - 1);
    t.x[n - 1 - i] = (double)z; t.w[n - 1 - i] = (double)(2 / ((1 - z * z) * pp * pp));
  }
  for (int i = 0; i < n / 2; i++) {
    double a = 0.5 * (t.x[n - 1 - i] - t.x[i]); t.x[i] = -a; t.x[n - 1 - i] = a;
    double b = 0.5 * (t.w[i] + t.w[n - 1 - i]); t.w[i] = b; t.w[n - 1 - i] = b;
  }
  if (n & 1) t.x[n / 2] = 0.0;
  return t;
}
