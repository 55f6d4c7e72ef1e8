// tma_probe.cu -- stand-alone probe of the TMA tile load used by the stencil (run each variant in its own process: a fault kills the context)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o tma_probe tma_probe.cu ; for v in 1 2 3 4 5 6; do ./tma_probe $v; done
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int RANK>
__global__ void k_probe(const __grid_constant__ CUtensorMap map, int c0, int c1, int c2, int bytes, float *out, int n, int fence_kind) {
  __shared__ __align__(128) float tile[4096];
  __shared__ __align__(8) unsigned long long bar;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
    if (fence_kind == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    else asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) tile[i] = -7.0f;
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(bytes) : "memory");
    if (RANK == 2)
      asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_u32(tile)), "l"(&map), "r"(c0), "r"(c1),
                   "r"(smem_u32(&bar))
                   : "memory");
    else
      asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(tile)), "l"(&map), "r"(c0),
                   "r"(c1), "r"(c2), "r"(smem_u32(&bar))
                   : "memory");
  }
  asm volatile("{\n .reg .pred p;\n WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE;\n bra WAIT;\n DONE:\n}" ::"r"(smem_u32(&bar)), "r"(0) : "memory");
  for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = tile[i];
}
typedef CUresult (*enc_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                           CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main(int argc, char **argv) {
  const int v = argc > 1 ? atoi(argv[1]) : 1;
  void *fn = nullptr; cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) { printf("v%d: no entry point\n", v); return 1; }
  enc_fn enc = (enc_fn)fn;
  const int nx = 519, ny = 519, nz = 8, px = 520;
  std::vector<float> h((size_t)px * ny * nz);
  for (int k = 0; k < nz; k++) for (int j = 0; j < ny; j++) for (int i = 0; i < px; i++) h[((size_t)k * ny + j) * px + i] = i < nx ? (float)(i + 1000 * j + 1000000 * k) : -1.0f;
  float *d, *out; cudaMalloc(&d, h.size() * 4); cudaMalloc(&out, 4096 * 4);
  cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
  CUtensorMap map; CUresult rc;
  int bx = 36, by = 20, c0 = -2, c1 = -2, c2 = -1, rank = 3, fence = 0;
  CUtensorMapL2promotion l2 = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;
  if (v == 1) { rank = 2; c0 = 0; c1 = 0; }
  if (v == 2) { rank = 2; }
  if (v == 3) { c2 = 1; c0 = 0; c1 = 0; }
  if (v == 4) { }                                   // what the stencil does
  if (v == 5) { l2 = CU_TENSOR_MAP_L2_PROMOTION_NONE; fence = 1; }
  if (v == 6) { bx = 32; by = 16; }
  if (v == 7) { rank = 2; c0 = -4; c1 = 0; }
  if (v == 8) { rank = 2; c0 = 0; c1 = -2; }
  if (v == 9) { bx = 40; c0 = -4; c1 = -2; c2 = -1; }
  if (v == 10) { rank = 2; c0 = 2; c1 = 0; }
  if (v == 11) { rank = 2; c0 = 4; c1 = 3; }
  if (v == 12) { bx = 40; c0 = 492; c1 = 510; c2 = 7; }      // overhanging the upper end in x and y
  if (v == 13) { bx = 40; c0 = -4; c1 = -2; c2 = 9; }        // a plane entirely outside
  cuuint64_t gdim[3] = {(cuuint64_t)nx, (cuuint64_t)ny, (cuuint64_t)nz}, gstr[2] = {(cuuint64_t)px * 4, (cuuint64_t)px * ny * 4};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
  rc = enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) { printf("v%d: encode failed %d\n", v, (int)rc); return 1; }
  const int n = bx * by;
  if (rank == 2) k_probe<2><<<1, 128>>>(map, c0, c1, c2, n * 4, out, n, fence); else k_probe<3><<<1, 128>>>(map, c0, c1, c2, n * 4, out, n, fence);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("v%d: FAULT %s\n", v, cudaGetErrorString(e)); return 2; }
  std::vector<float> r(n); cudaMemcpy(r.data(), out, n * 4, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int j = 0; j < by; j++) for (int i = 0; i < bx; i++) {
    const int gx = c0 + i, gy = c1 + j, gz = rank == 3 ? c2 : 0;
    const float want = (gx >= 0 && gx < nx && gy >= 0 && gy < ny && gz >= 0 && gz < nz) ? (float)(gx + 1000 * gy + 1000000 * gz) : 0.0f;
    if (r[j * bx + i] != want) bad++;
  }
  printf("v%d: ok, mismatches %d of %d (first %g %g %g)\n", v, bad, n, r[0], r[1], r[bx * 2 + 2]);
  return 0;
}
