// Throughput of packed FP32 FMA (fma.rn.f32x2 -> FFMA2) against scalar FFMA on sm_100a: does FFMA2 double the FP32 rate or only halve
// the issue slots?  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_probe ffma2_probe.cu && ./ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
  unsigned long long xa = *reinterpret_cast<unsigned long long *>(&a), ya = *reinterpret_cast<unsigned long long *>(&b), za = *reinterpret_cast<unsigned long long *>(&c), r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(xa), "l"(ya), "l"(za));
  return *reinterpret_cast<float2 *>(&r);
}
template <int MODE>      // 0: 16 scalar FFMA chains; 1: 8 FFMA2 chains (same flops); 2: 8 FFMA2 + 8 independent integer ops per round (issue pressure)
__global__ void __launch_bounds__(256) k(float *out, int iters, float a, float b) {
  float s[16]; int z = threadIdx.x;
#pragma unroll
  for (int q = 0; q < 16; q++) s[q] = threadIdx.x * 1e-3f + q;
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) {
#pragma unroll
      for (int q = 0; q < 16; q++) s[q] = fmaf(s[q], a, b);
    } else {
#pragma unroll
      for (int q = 0; q < 8; q++) { float2 r = ffma2(make_float2(s[2 * q], s[2 * q + 1]), make_float2(a, a), make_float2(b, b)); s[2 * q] = r.x; s[2 * q + 1] = r.y; }
      if (MODE == 2) {
#pragma unroll
        for (int q = 0; q < 8; q++) z = (z ^ (z >> 3)) + q;
      }
    }
  }
  float t = 0; for (int q = 0; q < 16; q++) t += s[q];
  out[blockIdx.x * blockDim.x + threadIdx.x] = t + z;
}
template <int MODE> void run(const char *name) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sms * 8, iters = 1 << 14; float *out; cudaMalloc(&out, sizeof(float) * blocks * 256);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 1e-7f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<blocks, 256>>>(out, iters, 1.0001f, 1e-7f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  printf("%-28s %.3f ms  %.1f TFLOP/s (FP32 flops only)  err=%s\n", name, ms, 2.0 * 16 * iters * (double)blocks * 256 / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out);
}
int main() { run<0>("16 x FFMA"); run<1>("8 x FFMA2"); run<2>("8 x FFMA2 + 16 int ops"); return 0; }
