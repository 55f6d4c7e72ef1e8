/* abi_smoke.c -- drives the C ABI of libr2s.so WITHOUT Python: what a Julia / C / Fortran host sees.
 * Builds a small structured HEX8 mesh with a spherical density, runs the pre-timer stages and the timed region of rho2sdf()
 * (RhoToSDF.jl:116-242) through r2s_pipeline and, when asked, through the single-call multi-GPU entry r2s_multi_pipeline
 * (slabs may share a device), and checks the two against each other.
 *   cc -std=c11 -I include tests/host/abi_smoke.c -o abi_smoke -L rho2sdf.jl_b200 -lr2s -Wl,-rpath,$PWD/rho2sdf.jl_b200 -lm
 *   ./abi_smoke [n=12] [nslabs=3]                exit code 0 = all checks passed                                                  */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "r2s.h"

#define CHECK(call) do { if ((call) != 0) { fprintf(stderr, "FAILED %s:%d: %s -> %s\n", __FILE__, __LINE__, #call, r2s_last_error(ctx)); return 1; } } while (0)
#define MCHECK(call) do { if ((call) != 0) { fprintf(stderr, "FAILED %s:%d: %s -> %s\n", __FILE__, __LINE__, #call, r2s_multi_last_error(m)); return 1; } } while (0)

int main(int argc, char **argv) {
  const int n = argc > 1 ? atoi(argv[1]) : 12, nslabs = argc > 2 ? atoi(argv[2]) : 3;
  const int64_t m1 = n + 1, nnp = m1 * m1 * m1, nel = (int64_t)n * n * n;
  double *X = malloc(sizeof(double) * 3 * nnp), *rho = malloc(sizeof(double) * nel), *rho_n = malloc(sizeof(double) * nnp);
  int64_t *IEN = malloc(sizeof(int64_t) * 8 * nel);
  for (int64_t k = 0; k < m1; k++) for (int64_t j = 0; j < m1; j++) for (int64_t i = 0; i < m1; i++) {
    const int64_t a = (k * m1 + j) * m1 + i; X[3 * a] = (double)i; X[3 * a + 1] = (double)j; X[3 * a + 2] = (double)k;      /* Julia: X is 3 x nnp */
  }
  for (int64_t k = 0; k < n; k++) for (int64_t j = 0; j < n; j++) for (int64_t i = 0; i < n; i++) {
    const int64_t e = (k * n + j) * n + i;
#define NID(a, b, c) (((c) * m1 + (b)) * m1 + (a) + 1)      /* 1-based, VTK hexahedron order */
    const int64_t v[8] = {NID(i, j, k), NID(i + 1, j, k), NID(i + 1, j + 1, k), NID(i, j + 1, k), NID(i, j, k + 1), NID(i + 1, j, k + 1), NID(i + 1, j + 1, k + 1), NID(i, j + 1, k + 1)};
    memcpy(IEN + 8 * e, v, sizeof(v));
    const double cx = i + 0.5 - n / 2.0, cy = j + 0.5 - n / 2.0, cz = k + 0.5 - n / 2.0, r = sqrt(cx * cx + cy * cy + cz * cz) / (0.5 * n);
    rho[e] = r < 0.55 ? 1.0 : (r < 0.8 ? 1.0 - (r - 0.55) / 0.25 : 0.0);
  }
  /* Grid ctor (Grid.jl:10-34) as the host computes it: cell = extent / N_max, 3 margin cells */
  const int N_max = 2 * n; const double cell = (double)n / N_max;
  double amin[3], amax[3]; int64_t N[3];
  for (int d = 0; d < 3; d++) { amin[d] = 0.0 - 3 * cell; amax[d] = (double)n + 3 * cell; N[d] = (int64_t)ceil((amax[d] - amin[d]) / cell); amax[d] = amin[d] + N[d] * cell; }
  const int64_t ngp = (N[0] + 1) * (N[1] + 1) * (N[2] + 1), nf = (2 * N[0] + 1) * (2 * N[1] + 1) * (2 * N[2] + 1);

  r2s_ctx *ctx = NULL;
  if (r2s_create(&ctx, 0, NULL) != 0) { fprintf(stderr, "r2s_create failed: no CUDA device (there is no CPU fallback)\n"); return 2; }
  double Vd, Vf, rho_t;
  CHECK(r2s_set_mesh(ctx, 8, nnp, X, nel, IEN));
  CHECK(r2s_mesh_volume(ctx, rho, &Vd, &Vf));
  CHECK(r2s_nodal_densities(ctx, rho, rho_n));
  CHECK(r2s_find_threshold(ctx, rho_n, Vd * Vf, 1e-4, 60, &rho_t));
  CHECK(r2s_set_grid(ctx, amin, amax, N, cell));
  r2s_params p; r2s_default_params(&p);
  p.rho_t = rho_t; p.smooth = 2; p.rbf_interp = 1; p.target_volume = Vd * Vf;
  double *sdf = malloc(sizeof(double) * ngp), *sdf2 = malloc(sizeof(double) * ngp);
  float *fine = malloc(sizeof(float) * nf), *fine2 = malloc(sizeof(float) * nf);
  r2s_report rep, rep2;
  CHECK(r2s_pipeline(ctx, &p, rho_n, sdf, fine, &rep));
  int64_t inside = 0, band = 0;
  for (int64_t v = 0; v < ngp; v++) { if (sdf[v] > 0) inside++; if (fabs(sdf[v]) < 1e9) band++; }
  printf("single: V_domain %.3f V_frac %.4f rho_t %.6f  pairs %lld (pruned %lld) cg %d th %.6f volume %.3f (target %.3f)  inside %lld band %lld launches %lld  %.2f ms\n",
         Vd, Vf, rho_t, (long long)rep.n_pairs, (long long)rep.n_pairs_pruned, rep.cg_iters, rep.th, rep.volume, Vd * Vf, (long long)inside, (long long)band, (long long)rep.launches, rep.ms_total);
  int bad = 0;
  if (!(fabs(Vd - (double)nel) < 1e-9 * nel)) { fprintf(stderr, "V_domain\n"); bad++; }
  if (!(rep.n_crossing > 0 && rep.cg_iters > 3 && inside > 0 && band > inside && rep.launches > 0)) { fprintf(stderr, "report\n"); bad++; }
  if (!(fabs(rep.volume - Vd * Vf) < 0.05 * Vd * Vf)) { fprintf(stderr, "volume\n"); bad++; }
  /* error behaviour: status + message, no exception crosses the ABI */
  if (r2s_set_mesh(ctx, 5, nnp, X, nel, IEN) == 0 || strstr(r2s_last_error(ctx), "nen") == NULL) { fprintf(stderr, "error path\n"); bad++; }

  if (nslabs > 1) {
    int dev[64]; for (int r = 0; r < nslabs; r++) dev[r] = 0;
    r2s_multi *m = NULL;
    if (r2s_multi_create(&m, dev, nslabs) != 0) { fprintf(stderr, "r2s_multi_create failed\n"); return 3; }
    MCHECK(r2s_multi_set_mesh(m, 8, nnp, X, nel, IEN));
    MCHECK(r2s_multi_set_grid(m, amin, amax, N, cell));
    for (int call = 0; call < 2; call++) MCHECK(r2s_multi_pipeline(m, &p, rho_n, sdf2, fine2, &rep2));      /* the second call runs on the re-cut slabs */
    double dmax = 0, fmax_ = 0;
    for (int64_t v = 0; v < ngp; v++) { double d = fabs(sdf[v] - sdf2[v]); if (d > dmax) dmax = d; }
    for (int64_t v = 0; v < nf; v++) { double d = fabs((double)fine[v] - (double)fine2[v]); if (d > fmax_) fmax_ = d; }
    printf("multi (%d slabs): max |sdf - single| %.3e  max |fine - single| %.3e  cg %d collectives %lld\n", nslabs, dmax, fmax_, rep2.cg_iters, (long long)rep2.collectives);
    if (dmax != 0.0 || fmax_ > 1e-4 || rep2.cg_iters != rep.cg_iters || rep2.collectives <= 0) { fprintf(stderr, "multi parity\n"); bad++; }
    r2s_multi_destroy(m);
  }
  r2s_destroy(ctx);
  printf(bad ? "ABI SMOKE FAILED (%d)\n" : "ABI SMOKE OK\n", bad);
  return bad ? 1 : 0;
}
