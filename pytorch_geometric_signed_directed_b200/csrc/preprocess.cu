// Sparse GPU preprocessing primitives: what the reference does with DENSE N x N
// matrices on the CPU before a DiGCN / DGCN model can run --
//   utils/directed/get_adjs_DiGCN.py:113-190  get_appr_directed_adj  (dense (N+1)^2 eig for the PPR
//                                              stationary vector, dense products with diag(pi^+-1/2))
//   utils/directed/get_adjs_DiGCN.py:193-254  get_second_directed_adj (dense P^T P and P P^T)
//   utils/directed/features_in_out.py:44-46   directed_features_in_out (N rank-1 sparse updates)
// -- is expressed here with three sparse primitives; the composition lives in
// pytorch_geometric_signed_directed_b200/utils/directed.py:
//   pgsd_coo_coalesce    sort (row * n + col) keys, sum duplicates          (CUB radix sort + reduce-by-key)
//   pgsd_gram_expand     all products of C = B^T diag(s) B, row by row      (expand step of an ESC SpGEMM)
//   pgsd_ppr_stationary  fp64 power iteration for the PPR stationary vector (sparse, replaces the dense eig)
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_reduce.cuh>

#include "common.cuh"

namespace pgsd {
namespace prep {

struct CoalesceSpace {
  uint64_t *keys_a, *keys_b;
  float *vals_a, *vals_b;
  int* n_unique;
  void* cub_temp;
  size_t cub_bytes, total;
};

static size_t cub_need(int64_t m) {
  size_t a = 0, b = 0;
  cub::DoubleBuffer<uint64_t> dk(nullptr, nullptr);
  cub::DoubleBuffer<float> dv(nullptr, nullptr);
  cub::DeviceRadixSort::SortPairs(nullptr, a, dk, dv, int(m), 0, 64, nullptr);
  cub::DeviceReduce::ReduceByKey(nullptr, b, static_cast<uint64_t*>(nullptr), static_cast<uint64_t*>(nullptr),
                                 static_cast<float*>(nullptr), static_cast<float*>(nullptr),
                                 static_cast<int*>(nullptr), cub::Sum(), int(m), nullptr);
  return (a > b ? a : b) + 256;
}

static CoalesceSpace carve(void* base, int64_t m) {
  CoalesceSpace s{};
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += align_up(bytes);
    return p;
  };
  s.keys_a = static_cast<uint64_t*>(take(size_t(m) * 8));
  s.keys_b = static_cast<uint64_t*>(take(size_t(m) * 8));
  s.vals_a = static_cast<float*>(take(size_t(m) * 4));
  s.vals_b = static_cast<float*>(take(size_t(m) * 4));
  s.n_unique = static_cast<int*>(take(256));
  s.cub_bytes = cub_need(m);
  s.cub_temp = take(s.cub_bytes);
  s.total = off;
  return s;
}

__global__ void k_copy_in(const int64_t* __restrict__ keys, const float* __restrict__ vals, int64_t m,
                          uint64_t* __restrict__ ka, float* __restrict__ va) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += int64_t(gridDim.x) * blockDim.x) {
    ka[i] = uint64_t(keys[i]);
    va[i] = vals[i];
  }
}
__global__ void k_copy_out(const uint64_t* __restrict__ k, int64_t m, int64_t* __restrict__ out) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < m; i += int64_t(gridDim.x) * blockDim.x)
    out[i] = int64_t(k[i]);
}

// One thread per product of C = B^T diag(s) B.  offs[k] = exclusive prefix of len_k^2; product t of row k
// pairs entry a = t / len with entry b = t % len: key = col[a] * n_cols + col[b], value = s_k val[a] val[b].
// Row k's products are contiguous and rows ascend, so after the stable radix sort the contributions to one
// (i, j) are summed in ascending k.
__global__ void k_gram_expand(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                              const float* __restrict__ val, const float* __restrict__ scale,
                              const int64_t* __restrict__ offs, int64_t n_rows, int64_t n_cols, int64_t total,
                              int64_t* __restrict__ keys, float* __restrict__ vals) {
  for (int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; t < total; t += int64_t(gridDim.x) * blockDim.x) {
    int64_t lo = 0, hi = n_rows;                       // largest k with offs[k] <= t
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (offs[mid] <= t) lo = mid; else hi = mid;
    }
    const int32_t s0 = row_ptr[lo], len = row_ptr[lo + 1] - s0;
    const int64_t r = t - offs[lo];
    const int32_t a = s0 + int32_t(r / len), b = s0 + int32_t(r % len);
    const float sc = scale ? scale[lo] : 1.f;
    keys[t] = int64_t(col[a]) * n_cols + col[b];
    vals[t] = (val ? val[a] * val[b] : 1.f) * sc;
  }
}

// y[j] = damp * sum_{e in row j} p[e] * x[col[e]] + teleport   (fp64 accumulate, fp32 transition weights)
__global__ void k_ppr_step(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                           const float* __restrict__ p, const double* __restrict__ x, double* __restrict__ y,
                           int64_t n, double damp, double teleport) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = int64_t(gridDim.x) * (blockDim.x >> 5);
  for (int64_t j = int64_t(blockIdx.x) * (blockDim.x >> 5) + (threadIdx.x >> 5); j < n; j += warps) {
    double acc = 0.0;
    for (int e = row_ptr[j] + lane; e < row_ptr[j + 1]; e += 32) acc += double(p[e]) * x[col[e]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) y[j] = damp * acc + teleport;
  }
}
__global__ void k_fill(double* x, int64_t n, double v) {
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += int64_t(gridDim.x) * blockDim.x) x[i] = v;
}

static int grid_for(int64_t n, int threads = 256) {
  int64_t g = ceil_div<int64_t>(n > 0 ? n : 1, threads);
  const int64_t cap = int64_t(sm_count()) * 16;
  return int(g > cap ? cap : g);
}

}  // namespace prep
}  // namespace pgsd

using namespace pgsd;
using namespace pgsd::prep;

extern "C" int pgsd_coalesce_workspace_bytes(int64_t n_entries, size_t* bytes_host) {
  PGSD_REQUIRE(bytes_host != nullptr, "coalesce_workspace_bytes: null output");
  PGSD_REQUIRE(n_entries >= 0 && n_entries < (int64_t(1) << 31), "coalesce: entry count exceeds the int32 range of the sort");
  *bytes_host = carve(nullptr, n_entries > 0 ? n_entries : 1).total + 256;
  return PGSD_OK;
}

extern "C" int pgsd_coo_coalesce(const int64_t* keys, const float* vals, int64_t n_entries, int key_bits,
                                 int64_t* keys_out, float* vals_out, int64_t* n_unique_host, void* workspace,
                                 size_t workspace_bytes, pgsd_stream_t stream) {
  PGSD_REQUIRE(n_unique_host != nullptr, "coo_coalesce: null count output");
  *n_unique_host = 0;
  if (n_entries == 0) return PGSD_OK;
  PGSD_REQUIRE(n_entries > 0 && n_entries < (int64_t(1) << 31), "coo_coalesce: entry count exceeds the int32 range of the sort");
  PGSD_REQUIRE(keys && vals && keys_out && vals_out && workspace, "coo_coalesce: null pointer");
  PGSD_REQUIRE(key_bits > 0 && key_bits <= 63, "coo_coalesce: key_bits must be in 1..63");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CoalesceSpace s = carve(workspace, n_entries);
  if (s.total > workspace_bytes) return fail(PGSD_ERR_WORKSPACE, "coo_coalesce: workspace %zu < %zu", workspace_bytes, s.total);
  k_copy_in<<<grid_for(n_entries), 256, 0, st>>>(keys, vals, n_entries, s.keys_a, s.vals_a);
  PGSD_LAUNCH_CHECK("k_copy_in");
  cub::DoubleBuffer<uint64_t> dk(s.keys_a, s.keys_b);
  cub::DoubleBuffer<float> dv(s.vals_a, s.vals_b);
  size_t need = s.cub_bytes;
  PGSD_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, need, dk, dv, int(n_entries), 0, key_bits, st));
  uint64_t* sorted_k = dk.Current();
  float* sorted_v = dv.Current();
  uint64_t* uniq_k = dk.Alternate();                   // the other halves are free now
  need = s.cub_bytes;
  PGSD_CUDA(cub::DeviceReduce::ReduceByKey(s.cub_temp, need, sorted_k, uniq_k, sorted_v, vals_out, s.n_unique,
                                           cub::Sum(), int(n_entries), st));
  int n_u = 0;
  PGSD_CUDA(cudaMemcpyAsync(&n_u, s.n_unique, sizeof(int), cudaMemcpyDeviceToHost, st));
  PGSD_CUDA(cudaStreamSynchronize(st));
  k_copy_out<<<grid_for(n_u), 256, 0, st>>>(uniq_k, n_u, keys_out);
  PGSD_LAUNCH_CHECK("k_copy_out");
  PGSD_CUDA(cudaStreamSynchronize(st));               // keys_out is read back from the workspace
  *n_unique_host = n_u;
  return PGSD_OK;
}

extern "C" int pgsd_gram_expand(const int32_t* row_ptr, const int32_t* col, const float* val, const float* scale,
                                const int64_t* product_offsets, int64_t n_rows, int64_t n_cols,
                                int64_t n_products, int64_t* keys_out, float* vals_out, pgsd_stream_t stream) {
  if (n_products == 0 || n_rows == 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr && col && product_offsets && keys_out && vals_out, "gram_expand: null pointer");
  PGSD_REQUIRE(n_cols > 0 && n_cols < (int64_t(1) << 31), "gram_expand: n_cols out of range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  k_gram_expand<<<grid_for(n_products), 256, 0, st>>>(row_ptr, col, val, scale, product_offsets, n_rows, n_cols,
                                                      n_products, keys_out, vals_out);
  PGSD_LAUNCH_CHECK("k_gram_expand");
  return PGSD_OK;
}

extern "C" int pgsd_ppr_stationary(const int32_t* row_ptr_dst, const int32_t* col_src, const float* p,
                                   int64_t n, double alpha, int32_t n_iter, double* pi, double* scratch,
                                   pgsd_stream_t stream) {
  if (n == 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr_dst && col_src && p && pi && scratch, "ppr_stationary: null pointer");
  PGSD_REQUIRE(alpha > 0.0 && alpha < 1.0 && n_iter > 0, "ppr_stationary: need 0 < alpha < 1 and n_iter > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // Stationary vector of the (N+1)-state chain of get_appr_directed_adj (get_adjs_DiGCN.py:147-152): with
  // S = sum(pi[0:N]) and t = pi[N], S' = (1-alpha) S + t and t' = alpha S, so at the fixed point (S + t = 1)
  // t = alpha / (1 + alpha) and pi[0:N] solves pi = (1 - alpha) P^T pi + t / N.  The iteration contracts by
  // (1 - alpha) per step; the caller picks n_iter from the accuracy it wants.
  const double teleport = alpha / (1.0 + alpha) / double(n);
  k_fill<<<grid_for(n), 256, 0, st>>>(pi, n, 1.0 / (1.0 + alpha) / double(n));
  PGSD_LAUNCH_CHECK("k_fill");
  double *src = pi, *dst = scratch;
  const int grid = grid_for(n * 32);
  for (int it = 0; it < n_iter; ++it) {
    k_ppr_step<<<grid, 256, 0, st>>>(row_ptr_dst, col_src, p, src, dst, n, 1.0 - alpha, teleport);
    double* t = src; src = dst; dst = t;
  }
  PGSD_LAUNCH_CHECK("k_ppr_step");
  if (src != pi) PGSD_CUDA(cudaMemcpyAsync(pi, src, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, st));
  return PGSD_OK;
}
