// Dense feature transform that sits next to every aggregation:
//   acc_g = sum_{t in group g} X_t @ W_t ;  y0 = acc_0 (-acc_1) + b ;  y1 = acc_0 + acc_1 + b
// Replaces the torch.matmul / nn.Linear call sites at nn/directed/MagNetConv.py:189-247,
// nn/directed/DiGCNConv.py:66, nn/directed/DiGCN_Inception_Block.py:44,
// nn/signed/SGCNConv.py:102-121 (the torch.cat of :102,105,113,120 is never materialised: each
// column block of the concatenation is one term).
//
// This file holds the exact-fp32 FFMA path (round-1 baseline for the transform): a 64x64
// output tile per CTA, 4x4 register block per thread, K streamed through shared memory in
// 32-wide slabs.  Arithmetic is plain fp32 FMA, i.e. the same class as the reference's
// MKL/cuBLAS sgemm with allow_tf32 = False.
#include "common.cuh"

namespace pgsd {

constexpr int BM = 64, BN = 64, BK = 32, DT = 256;

struct DenseParams {
  int64_t n_rows;
  int32_t n_out, n_terms, combine, relu_mode;
  const char* x[PGSD_DENSE_MAX_TERMS];
  int64_t ldx[PGSD_DENSE_MAX_TERMS];  // elements
  int32_t k[PGSD_DENSE_MAX_TERMS];
  int32_t group[PGSD_DENSE_MAX_TERMS];
  const float* w[PGSD_DENSE_MAX_TERMS];
  int64_t ldw_k[PGSD_DENSE_MAX_TERMS], ldw_n[PGSD_DENSE_MAX_TERMS];
  const float* bias;
  char* y[2];
  int64_t ldy[2];
};

template <bool BF16>
__device__ __forceinline__ float load_x(const char* base, int64_t idx) {
  if constexpr (BF16)
    return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(base)[idx]);
  else
    return __ldg(reinterpret_cast<const float*>(base) + idx);
}

// GROUPS = 1: single accumulator.  GROUPS = 2: MagNet mixing; terms alternate between the two
// accumulators but share W when they come in (real, imag) pairs -- here each term simply
// carries its own W pointer.
template <bool BF16, int GROUPS>
__global__ void __launch_bounds__(DT) dense_ffma_kernel(const DenseParams p) {
  __shared__ __align__(16) float Xs[BK][BM + 4];
  __shared__ __align__(16) float Ws[BK][BN + 4];

  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t row0 = int64_t(blockIdx.x) * BM;
  const int col0 = blockIdx.y * BN;

  float acc[GROUPS][4][4];
#pragma unroll
  for (int g = 0; g < GROUPS; ++g)
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[g][i][j] = 0.f;

  const int xr = tid >> 2;        // 0..63: row inside the tile this thread stages
  const int xk = (tid & 3) * 8;   // 8 consecutive k per thread
  const int wk = tid >> 3;        // 0..31
  const int wn = (tid & 7) * 8;   // 8 consecutive n per thread

  for (int t = 0; t < p.n_terms; ++t) {
    const int kt = p.k[t];
    const int grp = GROUPS == 2 ? p.group[t] : 0;
    const char* xb = p.x[t];
    const int64_t ldx = p.ldx[t];
    const float* wb = p.w[t];
    const int64_t lwk = p.ldw_k[t], lwn = p.ldw_n[t];
    for (int k0 = 0; k0 < kt; k0 += BK) {
      // stage X^T and W slabs (zero padded)
      {
        const int64_t r = row0 + xr;
        const bool rok = r < p.n_rows;
        const bool fast = !BF16 && rok && (k0 + xk + 8 <= kt) && ((ldx & 3) == 0) &&
                          ((reinterpret_cast<uintptr_t>(xb) & 15) == 0);
        if (fast) {
          const float4* q = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(xb) +
                                                            r * ldx + k0 + xk);
          const float4 a = __ldg(q), b = __ldg(q + 1);
          Xs[xk + 0][xr] = a.x, Xs[xk + 1][xr] = a.y, Xs[xk + 2][xr] = a.z, Xs[xk + 3][xr] = a.w;
          Xs[xk + 4][xr] = b.x, Xs[xk + 5][xr] = b.y, Xs[xk + 6][xr] = b.z, Xs[xk + 7][xr] = b.w;
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int kk = k0 + xk + i;
            Xs[xk + i][xr] = (rok && kk < kt) ? load_x<BF16>(xb, r * ldx + kk) : 0.f;
          }
        }
        const int kk = k0 + wk;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int n = col0 + wn + j;
          Ws[wk][wn + j] = (kk < kt && n < p.n_out) ? __ldg(wb + kk * lwk + n * lwn) : 0.f;
        }
      }
      __syncthreads();
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        const float4 a = *reinterpret_cast<const float4*>(&Xs[kk][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
        if (GROUPS == 1 || grp == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[0][i][j] = fmaf(av[i], bv[j], acc[0][i][j]);
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
              acc[GROUPS - 1][i][j] = fmaf(av[i], bv[j], acc[GROUPS - 1][i][j]);
        }
      }
      __syncthreads();
    }
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t r = row0 + ty * 4 + i;
    if (r >= p.n_rows) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = col0 + tx * 4 + j;
      if (n >= p.n_out) continue;
      const float b = p.bias ? __ldg(p.bias + n) : 0.f;
      float o0, o1 = 0.f;
      if (GROUPS == 2) {
        // out_real = A - B + b ; out_imag = A + B + b      (MagNetConv.py:242-247)
        o0 = (acc[0][i][j] - acc[GROUPS - 1][i][j]) + b;
        o1 = (acc[0][i][j] + acc[GROUPS - 1][i][j]) + b;
        if (p.relu_mode == 1) {  // complex_relu.py:21-22: mask = 1.0 * (real >= 0)
          const float m = o0 >= 0.f ? 1.f : 0.f;
          o0 *= m, o1 *= m;
        }
      } else {
        o0 = acc[0][i][j] + b;
        if (p.relu_mode == 2) o0 = tanhf(o0);   // SGCN.py:93-96: z = tanh(conv(...))
      }
      if constexpr (BF16) {
        reinterpret_cast<__nv_bfloat16*>(p.y[0])[r * p.ldy[0] + n] = __float2bfloat16_rn(o0);
        if (GROUPS == 2)
          reinterpret_cast<__nv_bfloat16*>(p.y[1])[r * p.ldy[1] + n] = __float2bfloat16_rn(o1);
      } else {
        reinterpret_cast<float*>(p.y[0])[r * p.ldy[0] + n] = o0;
        if (GROUPS == 2) reinterpret_cast<float*>(p.y[1])[r * p.ldy[1] + n] = o1;
      }
    }
  }
}

int dense_tc_try(const pgsd_dense_args* a, cudaStream_t st, int* handled);  // dense_tc.cu
int dense_tma_try(const pgsd_dense_args* a, cudaStream_t st, int* handled); // dense_tma.cu

}  // namespace pgsd

using namespace pgsd;

extern "C" int pgsd_dense_transform(const pgsd_dense_args* a, pgsd_stream_t stream) {
  PGSD_REQUIRE(a != nullptr, "dense: args is null");
  PGSD_REQUIRE(a->n_terms >= 1 && a->n_terms <= PGSD_DENSE_MAX_TERMS, "dense: n_terms=%d out of range",
               a->n_terms);
  PGSD_REQUIRE(a->dtype == PGSD_F32 || a->dtype == PGSD_BF16, "dense: bad dtype");
  PGSD_REQUIRE(a->combine == 0 || a->combine == 1, "dense: bad combine");
  PGSD_REQUIRE(a->n_rows >= 0 && a->n_out >= 0, "dense: negative size");
  if (a->n_rows == 0 || a->n_out == 0) return PGSD_OK;
  PGSD_REQUIRE(a->y[0] && (a->combine == 0 || a->y[1]), "dense: null output");
  DenseParams p{};
  p.n_rows = a->n_rows;
  p.n_out = a->n_out;
  p.n_terms = a->n_terms;
  p.combine = a->combine;
  p.relu_mode = a->relu_mode;
  p.bias = a->bias;
  for (int t = 0; t < a->n_terms; ++t) {
    PGSD_REQUIRE(a->x[t] && a->w[t] && a->k[t] >= 0, "dense: term %d has a null pointer", t);
    PGSD_REQUIRE(a->group[t] == 0 || (a->combine == 1 && a->group[t] == 1), "dense: bad group");
    p.x[t] = static_cast<const char*>(a->x[t]);
    p.ldx[t] = a->ldx[t];
    p.k[t] = a->k[t];
    p.group[t] = a->group[t];
    p.w[t] = a->w[t];
    p.ldw_k[t] = a->ldw_k[t];
    p.ldw_n[t] = a->ldw_n[t];
  }
  for (int i = 0; i < 2; ++i) {
    p.y[i] = static_cast<char*>(a->y[i]);
    p.ldy[i] = a->ldy[i];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  for (int t = 0; t < a->n_terms; ++t) {
    PGSD_REQUIRE(a->x[t] && a->w[t] && a->k[t] >= 0, "dense: term %d has a null pointer", t);
    PGSD_REQUIRE(a->group[t] == 0 || (a->combine == 1 && a->group[t] == 1), "dense: bad group");
  }
  // variant: 0 = auto (TMA-fed tcgen05 kernel, else register-staged tcgen05 kernel, else FFMA), 1 = force FFMA,
  // 2 = require the tcgen05 path, 4 = require its warp-specialised kernel
  // 16 = require the TMA-fed warp-specialised kernel (dense_tma.cu); its ring depths may ride in bits 8-15
  if ((a->variant & 0xff) == 16) {
    int handled = 0;
    int rc = dense_tma_try(a, st, &handled);
    if (rc != PGSD_OK) return rc;
    if (handled) return PGSD_OK;
    return fail(PGSD_ERR_INVALID, "dense: shape outside the TMA kernel's envelope");
  }
  if (a->variant == 0) {      // auto: the TMA-fed kernel first (measured 1.2-1.4x the register-staged one)
    int handled = 0;
    int rc = dense_tma_try(a, st, &handled);
    if (rc != PGSD_OK) return rc;
    if (handled) return PGSD_OK;
  }
  if (a->variant != 1) {
    int handled = 0;
    int rc = dense_tc_try(a, st, &handled);
    if (rc != PGSD_OK) return rc;
    if (handled) return PGSD_OK;
    if (a->variant == 2 || a->variant == 4 || a->variant == 8) return fail(PGSD_ERR_INVALID, "dense: shape outside the tcgen05 path's envelope");
  }
  dim3 grid((unsigned)ceil_div<int64_t>(a->n_rows, BM), (unsigned)ceil_div<int>(a->n_out, BN));
  if (a->dtype == PGSD_BF16) {
    if (a->combine) dense_ffma_kernel<true, 2><<<grid, DT, 0, st>>>(p);
    else dense_ffma_kernel<true, 1><<<grid, DT, 0, st>>>(p);
  } else {
    if (a->combine) dense_ffma_kernel<false, 2><<<grid, DT, 0, st>>>(p);
    else dense_ffma_kernel<false, 1><<<grid, DT, 0, st>>>(p);
  }
  PGSD_LAUNCH_CHECK("dense_ffma_kernel");
  return PGSD_OK;
}

extern "C" int pgsd_abi_version(void) { return PGSD_ABI_VERSION; }
extern "C" int pgsd_sizeof_args(size_t* spmm_args_bytes_host, size_t* dense_args_bytes_host) {
  if (spmm_args_bytes_host) *spmm_args_bytes_host = sizeof(pgsd_spmm_args);
  if (dense_args_bytes_host) *dense_args_bytes_host = sizeof(pgsd_dense_args);
  return PGSD_OK;
}
extern "C" const char* pgsd_last_error(void) { return pgsd::err_buf(); }
extern "C" int pgsd_device_info(int* sm_count_host, int* cc_major_host, int* cc_minor_host) {
  int dev = 0;
  PGSD_CUDA(cudaGetDevice(&dev));
  if (sm_count_host) PGSD_CUDA(cudaDeviceGetAttribute(sm_count_host, cudaDevAttrMultiProcessorCount, dev));
  if (cc_major_host) PGSD_CUDA(cudaDeviceGetAttribute(cc_major_host, cudaDevAttrComputeCapabilityMajor, dev));
  if (cc_minor_host) PGSD_CUDA(cudaDeviceGetAttribute(cc_minor_host, cudaDevAttrComputeCapabilityMinor, dev));
  return PGSD_OK;
}
