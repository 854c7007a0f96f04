// Backward-pass helper of the dense transform (the reference trains every model
// through autograd of index_select / scatter_add_ / matmul):
//
//   dW[k, n]  += sum_r X[r, k] * G[r, n]        (weight gradient of y = X W, X: [R, K], G: [R, N])
//   db[n]     += sum_r G[r, n]                  (bias gradient, optional)
//
// Tall-skinny reduction over up to millions of rows: each CTA walks row slabs of 64 rows through
// shared memory, keeps a 4x4 register block of the [K, N] result per thread (64x64 tile per
// CTA.y/z), and adds its partial sums to the fp32 result with atomics at the end (R/64 slabs are
// folded inside the CTA first, so a result element sees at most gridDim.x atomic adds).
// The input gradients dX = G W^T and the transposed aggregation reuse pgsd_dense_transform /
// pgsd_spmm_csr.
#include "common.cuh"

namespace pgsd {

constexpr int GT = 256, GBR = 64, GBK = 64, GBN = 64;

template <bool BF16>
__device__ __forceinline__ float ldx(const char* base, int64_t idx) {
  if constexpr (BF16)
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  else
    return __ldg(reinterpret_cast<const float*>(base) + idx);
}

template <bool BF16>
__global__ void __launch_bounds__(GT) xtg_kernel(const char* __restrict__ x, int64_t ldx_,
                                                 const char* __restrict__ g, int64_t ldg_, int64_t n_rows,
                                                 int kdim, int ndim, float* __restrict__ dw, int64_t lddw,
                                                 float* __restrict__ db) {
  __shared__ __align__(16) float Xs[GBR][GBK + 4];
  __shared__ __align__(16) float Gs[GBR][GBN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int k0 = blockIdx.y * GBK, n0 = blockIdx.z * GBN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  float bsum[4] = {0.f, 0.f, 0.f, 0.f};
  const bool do_bias = db != nullptr && blockIdx.y == 0 && ty == 0;

  const int lr = tid >> 2, lc = (tid & 3) * 16;   // staging: row lr, 16 consecutive columns from lc
  for (int64_t r0 = int64_t(blockIdx.x) * GBR; r0 < n_rows; r0 += int64_t(gridDim.x) * GBR) {
    const int64_t r = r0 + lr;
#pragma unroll
    for (int c = 0; c < 16; ++c) {
      const int kk = k0 + lc + c, nn = n0 + lc + c;
      Xs[lr][lc + c] = (r < n_rows && kk < kdim) ? ldx<BF16>(x, r * ldx_ + kk) : 0.f;
      Gs[lr][lc + c] = (r < n_rows && nn < ndim) ? ldx<BF16>(g, r * ldg_ + nn) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int rr = 0; rr < GBR; ++rr) {
      const float4 a = *reinterpret_cast<const float4*>(&Xs[rr][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Gs[rr][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
      if (do_bias) {
#pragma unroll
        for (int j = 0; j < 4; ++j) bsum[j] += bv[j];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int kk = k0 + ty * 4 + i;
    if (kk >= kdim) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn < ndim) atomicAdd(dw + kk * lddw + nn, acc[i][j]);
    }
  }
  if (do_bias) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int nn = n0 + tx * 4 + j;
      if (nn < ndim) atomicAdd(db + nn, bsum[j]);
    }
  }
}

}  // namespace pgsd

using namespace pgsd;

extern "C" int pgsd_xtg_accumulate(const void* x, int64_t ldx, const void* g, int64_t ldg, int64_t n_rows,
                                   int32_t k, int32_t n, int32_t dtype, float* dw, int64_t lddw, float* db,
                                   pgsd_stream_t stream) {
  PGSD_REQUIRE(dtype == PGSD_F32 || dtype == PGSD_BF16, "xtg: bad dtype");
  PGSD_REQUIRE(n_rows >= 0 && k >= 0 && n >= 0, "xtg: negative size");
  if (n_rows == 0 || k == 0 || n == 0) return PGSD_OK;
  PGSD_REQUIRE(x && g && dw, "xtg: null pointer");
  int64_t gx = ceil_div<int64_t>(n_rows, GBR);
  const int64_t cap = int64_t(sm_count()) * 4;
  if (gx > cap) gx = cap;
  dim3 grid((unsigned)gx, (unsigned)ceil_div<int>(k, GBK), (unsigned)ceil_div<int>(n, GBN));
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == PGSD_BF16)
    xtg_kernel<true><<<grid, GT, 0, st>>>(static_cast<const char*>(x), ldx, static_cast<const char*>(g), ldg,
                                          n_rows, k, n, dw, lddw, db);
  else
    xtg_kernel<false><<<grid, GT, 0, st>>>(static_cast<const char*>(x), ldx, static_cast<const char*>(g), ldg,
                                           n_rows, k, n, dw, lddw, db);
  PGSD_LAUNCH_CHECK("xtg_kernel");
  return PGSD_OK;
}
