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

// ---- dL/dq of a trainable magnetic charge (SDDMM reduced to a scalar) ------------------------
// A group of LPR lanes owns a row: it keeps its slice of the two output-gradient rows in
// registers and walks the row's entries, gathering the matching slices of x_real / x_imag.
// No per-entry reduction is needed (the result is one scalar): every lane accumulates
// theta * (val_i * g_r.x_r - val_r * g_i.x_i) for its slice, lanes are summed once at the end.
namespace pgsd {

template <int VEC>
struct FVec;
template <>
struct FVec<4> {
  float4 v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(reinterpret_cast<const float4*>(p)); }
  __device__ __forceinline__ float dot(const FVec& o) const {
    return fmaf(v.x, o.v.x, fmaf(v.y, o.v.y, fmaf(v.z, o.v.z, v.w * o.v.w)));
  }
};
template <>
struct FVec<1> {
  float v;
  __device__ __forceinline__ void load(const float* p) { v = __ldg(p); }
  __device__ __forceinline__ float dot(const FVec& o) const { return v * o.v; }
};

constexpr int QG_THREADS = 256;

template <int VEC>
__global__ void __launch_bounds__(QG_THREADS) magnetic_q_grad_kernel(
    const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col, const float* __restrict__ vr,
    const float* __restrict__ vi, const float* __restrict__ th, int64_t n_rows, int feat,
    const float* __restrict__ gr, int64_t ldgr, const float* __restrict__ gi, int64_t ldgi,
    const float* __restrict__ xr, int64_t ldxr, const float* __restrict__ xi, int64_t ldxi, int lpr,
    double scale, double* __restrict__ dq) {
  const int lane = threadIdx.x & 31, sub = lane & (lpr - 1), rows_per_warp = 32 / lpr;
  const int64_t warp = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  float acc = 0.f;
  for (int64_t r0 = warp * rows_per_warp; r0 < n_rows; r0 += n_warps * rows_per_warp) {
    const int64_t r = r0 + lane / lpr;
    if (r >= n_rows) continue;
    const int beg = row_ptr[r], end = row_ptr[r + 1];
    for (int c = sub * VEC; c < feat; c += lpr * VEC) {
      FVec<VEC> g_r, g_i;
      g_r.load(gr + r * ldgr + c);
      g_i.load(gi + r * ldgi + c);
      int e = beg;
      for (; e + 4 <= end; e += 4) {
        int cc[4];
        float a[4], b[4];
        FVec<VEC> u[4], w[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          cc[j] = __ldg(col + e + j);
          const float t = __ldg(th + e + j);
          a[j] = t * __ldg(vi + e + j);
          b[j] = t * __ldg(vr + e + j);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          u[j].load(xr + int64_t(cc[j]) * ldxr + c);
          w[j].load(xi + int64_t(cc[j]) * ldxi + c);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) acc += a[j] * g_r.dot(u[j]) - b[j] * g_i.dot(w[j]);
      }
      for (; e < end; ++e) {
        const int cj = __ldg(col + e);
        const float t = __ldg(th + e);
        FVec<VEC> u, w;
        u.load(xr + int64_t(cj) * ldxr + c);
        w.load(xi + int64_t(cj) * ldxi + c);
        acc += t * __ldg(vi + e) * g_r.dot(u) - t * __ldg(vr + e) * g_i.dot(w);
      }
    }
  }
  __shared__ double part[QG_THREADS / 32];
  double d = double(acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) d += __shfl_xor_sync(0xffffffffu, d, o);
  if (lane == 0) part[threadIdx.x >> 5] = d;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < QG_THREADS / 32; ++k) s += part[k];
    atomicAdd(dq, s * scale);
  }
}

}  // namespace pgsd

extern "C" int pgsd_magnetic_q_grad(const int32_t* row_ptr, const int32_t* col, const float* val_real,
                                    const float* val_imag, const float* theta, int64_t n_rows, int32_t feat,
                                    const float* gy_real, int64_t ldgr, const float* gy_imag, int64_t ldgi,
                                    const float* x_real, int64_t ldxr, const float* x_imag, int64_t ldxi,
                                    double scale, double* dq, pgsd_stream_t stream) {
  PGSD_REQUIRE(n_rows >= 0 && feat >= 0, "magnetic_q_grad: negative size");
  PGSD_REQUIRE(dq != nullptr, "magnetic_q_grad: null dq");
  if (n_rows == 0 || feat == 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr && gy_real && gy_imag && x_real && x_imag, "magnetic_q_grad: null pointer");
  const bool vec4 = feat % 4 == 0 && ldgr % 4 == 0 && ldgi % 4 == 0 && ldxr % 4 == 0 && ldxi % 4 == 0 &&
                    ((reinterpret_cast<uintptr_t>(gy_real) | reinterpret_cast<uintptr_t>(gy_imag) |
                      reinterpret_cast<uintptr_t>(x_real) | reinterpret_cast<uintptr_t>(x_imag)) & 15) == 0;
  const int chunks = vec4 ? feat / 4 : feat;
  int lpr = 1;
  while (lpr < 32 && lpr < chunks) lpr <<= 1;
  const int64_t rows_per_block = int64_t(QG_THREADS / 32) * (32 / lpr);
  int64_t grid = ceil_div<int64_t>(n_rows, rows_per_block);
  const int64_t cap = int64_t(sm_count()) * 8;
  if (grid > cap) grid = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (vec4)
    magnetic_q_grad_kernel<4><<<unsigned(grid), QG_THREADS, 0, st>>>(
        row_ptr, col, val_real, val_imag, theta, n_rows, feat, gy_real, ldgr, gy_imag, ldgi, x_real, ldxr,
        x_imag, ldxi, lpr, scale, dq);
  else
    magnetic_q_grad_kernel<1><<<unsigned(grid), QG_THREADS, 0, st>>>(
        row_ptr, col, val_real, val_imag, theta, n_rows, feat, gy_real, ldgr, gy_imag, ldgi, x_real, ldxr,
        x_imag, ldxi, lpr, scale, dq);
  PGSD_LAUNCH_CHECK("magnetic_q_grad_kernel");
  return PGSD_OK;
}
