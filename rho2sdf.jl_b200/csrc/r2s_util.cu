// r2s_util.cu -- device-wide scan / radix sort plumbing (CUB from the CUDA toolkit) used by the binning steps
#include <cub/cub.cuh>
#include "r2s_common.cuh"

int r2s_scan_exclusive_i64(r2s_ctx *ctx, const i64 *in, i64 *out, i64 n) {
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream));
  CK(ctx->cubtmp.reserve(tmp));
  CK(cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, tmp, in, out, (int)n, ctx->stream));
  ctx->launches += 2;
  return 0;
}
int r2s_scan_exclusive_i32(r2s_ctx *ctx, const int *in, int *out, i64 n) {
  size_t tmp = 0;
  CK(cub::DeviceScan::ExclusiveSum(nullptr, tmp, in, out, (int)n, ctx->stream));
  CK(ctx->cubtmp.reserve(tmp));
  CK(cub::DeviceScan::ExclusiveSum(ctx->cubtmp.p, tmp, in, out, (int)n, ctx->stream));
  ctx->launches += 2;
  return 0;
}
// sorts n 64-bit keys on bits [0,end_bit); *sorted points at whichever of keys/alt holds the result
int r2s_sort_keys_u64(r2s_ctx *ctx, u64 *keys, u64 *alt, i64 n, int end_bit, u64 **sorted) {
  cub::DoubleBuffer<u64> db(keys, alt);
  size_t tmp = 0;
  CK(cub::DeviceRadixSort::SortKeys(nullptr, tmp, db, (int)n, 0, end_bit, ctx->stream));
  CK(ctx->cubtmp.reserve(tmp));
  CK(cub::DeviceRadixSort::SortKeys(ctx->cubtmp.p, tmp, db, (int)n, 0, end_bit, ctx->stream));
  ctx->launches += (end_bit + 7) / 8 + 1;
  *sorted = db.Current();
  return 0;
}
