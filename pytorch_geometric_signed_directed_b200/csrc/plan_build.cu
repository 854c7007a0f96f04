// Plan builders: user COO edge lists (int64, arbitrary order, duplicates, self-loops) ->
// CSR-by-destination plans (int32) consumed by spmm.cu.  Built once per graph and cached by
// the Python layers under the reference's cache rules, so none of this is on the
// steady-state path -- but with cached=False (the reference's default for MagNetConv) it runs
// inside every forward, where the reference spends ~25 s of CPU time on a 40M-key sort
// (SURVEY a3/a4).
//
// Replaces:
//   utils/directed/get_magnetic_Laplacian.py:44-87, utils/general/get_magnetic_signed_Laplacian.py:45-92
//   nn/directed/MagNetConv.py:78-120 (__norm__), nn/general/conv_base.py:12-31 (conv_norm_rw)
//   and the implicit index plumbing of PyG propagate.
//
// The only library primitives used are cub::DeviceRadixSort (stable LSD sort == PyG coalesce's
// stable index_sort) and cub::DeviceScan; CUB ships inside the CUDA toolkit.  Everything else
// (key construction, duplicate reduction in edge order, degree, phase, normalisation, row
// pointers) is written here.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "common.cuh"

namespace pgsd {

constexpr int TPB = 256;

struct Counters {
  int err;       // 1: node index out of range
  int n_loops;   // number of self-loop edges seen
  int nnz;       // stored entries
  int pad;
};

static inline int bit_length(uint64_t v) {
  int b = 0;
  while (v) ++b, v >>= 1;
  return b;
}
static inline unsigned blocks_for(int64_t n, int per = TPB) {
  int64_t b = ceil_div<int64_t>(n, per);
  const int64_t cap = int64_t(sm_count()) * 32;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return unsigned(b);
}
#define GRID_STRIDE(i, n)                                                        \
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < (n);      \
       i += int64_t(gridDim.x) * blockDim.x)

// ---------------------------------------------------------------------------------- kernels

// slots [0,E): (row,col) ; slots [E,2E): (col,row).  Self loops get the sentinel key, and so do
// entries whose row lies outside [row_lo, row_hi) (row-range build of one shard; rows are re-based).
__global__ void k_keys_sym(const int64_t* __restrict__ row, const int64_t* __restrict__ col,
                           int64_t E, int64_t N, int64_t row_lo, int64_t row_hi, uint64_t sentinel,
                           uint64_t* __restrict__ keys, uint32_t* __restrict__ slots, Counters* cnt) {
  GRID_STRIDE(s, 2 * E) {
    const int64_t p = s < E ? s : s - E;
    const int64_t r = row[p], c = col[p];
    uint64_t key = sentinel;
    if (r < 0 || r >= N || c < 0 || c >= N) {
      cnt->err = 1;
    } else if (r != c) {
      const int64_t a = s < E ? r : c, b = s < E ? c : r;
      if (a >= row_lo && a < row_hi) key = uint64_t(a - row_lo) * uint64_t(N) + uint64_t(b);
    }
    keys[s] = key;
    slots[s] = uint32_t(s);
  }
}

__global__ void k_heads64(const uint64_t* __restrict__ keys, int64_t M, uint64_t sentinel,
                          int32_t* __restrict__ head) {
  GRID_STRIDE(s, M) {
    const uint64_t k = keys[s];
    head[s] = (k != sentinel && (s == 0 || k != keys[s - 1])) ? 1 : 0;
  }
}

// One thread per unique (a,b) entry: sum its duplicates in sorted (= original slot) order,
// exactly the order PyG coalesce's scatter_add visits them.
__global__ void k_reduce_dups(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ slots,
                              const int32_t* __restrict__ head, const int32_t* __restrict__ uidx,
                              const float* __restrict__ weight, int64_t E, int64_t M, int64_t N,
                              int32_t* __restrict__ urow, int32_t* __restrict__ ucol,
                              float* __restrict__ usym, float* __restrict__ utheta,
                              float* __restrict__ uabs) {
  GRID_STRIDE(s, M) {
    if (!head[s]) continue;
    const uint64_t k = keys[s];
    float sym = 0.f, theta = 0.f, ab = 0.f;
    for (int64_t j = s; j < M && keys[j] == k; ++j) {
      const uint32_t slot = slots[j];
      const int64_t p = slot < E ? slot : slot - E;
      const float w = weight ? weight[p] : 1.f;
      sym += w;
      theta += slot < E ? w : -w;
      ab += fabsf(w);
    }
    const int32_t u = uidx[s];
    urow[u] = int32_t(k / uint64_t(N));
    ucol[u] = int32_t(k % uint64_t(N));
    usym[u] = sym;
    utheta[u] = theta;
    uabs[u] = ab;
  }
}

__global__ void k_set_nnz(const int32_t* head, const int32_t* uidx, int64_t M, Counters* cnt) {
  if (threadIdx.x == 0 && blockIdx.x == 0) cnt->nnz = uidx[M - 1] + head[M - 1];
}

// row_ptr from a sorted row-id array of (device-side) length *n_ptr.
__global__ void k_row_ptr(const int32_t* __restrict__ rows, const int* n_ptr, int64_t cap,
                          int64_t N, int32_t* __restrict__ row_ptr) {
  const int64_t n = *n_ptr;
  GRID_STRIDE(u, cap + 1) {
    if (u > n) continue;
    const int64_t prev = u == 0 ? -1 : rows[u - 1];
    const int64_t cur = u == n ? N : rows[u];
    for (int64_t r = prev + 1; r <= cur; ++r) row_ptr[r] = int32_t(u);
  }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// deg[r] = sum over row r of the symmetrised weight (or its absolute variant).
__global__ void k_degree(const int32_t* __restrict__ row_ptr, const float* __restrict__ usym,
                         const float* __restrict__ uabs, int signed_mode, int64_t N,
                         float* __restrict__ deg) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < N;
       r += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    float acc = 0.f;
    for (int e = row_ptr[r] + lane; e < row_ptr[r + 1]; e += 32) {
      float t;
      if (signed_mode == 0) t = usym[e] / 2;
      else if (signed_mode == 1) t = uabs[e] / 2;
      else t = fabsf(usym[e] / 2);
      acc += t;
    }
    acc = warp_sum(acc);
    if (lane == 0) deg[r] = acc;
  }
}

__device__ __forceinline__ float inv_sqrt_or_zero(float d) {
  const float t = 1.0f / sqrtf(d);  // deg.pow(-0.5); inf -> 0 (get_magnetic_Laplacian.py:76-77)
  return t == INFINITY ? 0.f : t;
}
__device__ __forceinline__ float scale_lmax(float v, float lambda_max) {
  const float t = (2.0f * v) / lambda_max;  // MagNetConv.py:106-107,115-116
  return t == INFINITY ? 0.f : t;
}

// Entry u of destination row a = urow[u], source b = ucol[u] carries L~[b,a]
// (source_to_target aggregation: out[a] += L~[b,a] * x[b]).  Theta(b,a) = -Theta(a,b).
__global__ void k_magnetic_values(const int32_t* __restrict__ urow, const int32_t* __restrict__ ucol,
                                  const float* __restrict__ usym, const float* __restrict__ utheta,
                                  const float* __restrict__ deg, const int* n_ptr, float two_pi_q,
                                  int normalization, float lambda_max,
                                  float* __restrict__ val_real, float* __restrict__ val_imag,
                                  float* __restrict__ theta_out) {
  const int64_t n = *n_ptr;
  GRID_STRIDE(u, n) {
    const int a = urow[u], b = ucol[u];
    if (theta_out) theta_out[u] = utheta[u];
    const float s = usym[u] / 2;
    float sn, cs;
    sincosf(two_pi_q * utheta[u], &sn, &cs);
    float mag;
    if (normalization) mag = (inv_sqrt_or_zero(deg[b]) * s) * inv_sqrt_or_zero(deg[a]);
    else mag = s;
    val_real[u] = scale_lmax(-(mag * cs), lambda_max);
    val_imag[u] = scale_lmax(mag * sn, lambda_max);
  }
}

// Row-range variant (second phase of the sharded build): local row r is global node row_lo + r and
// `deg` spans all nodes (all-gathered between the phases).  Same arithmetic, same operand order.
__global__ void k_magnetic_values_rows(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ ucol,
                                       const float* usym, const float* utheta,   // may alias the outputs
                                       const float* __restrict__ deg, int64_t n_rows, int64_t row_lo,
                                       float two_pi_q, int normalization, float lambda_max,
                                       float* val_real, float* val_imag) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < n_rows;
       r += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const float da = normalization ? inv_sqrt_or_zero(deg[row_lo + r]) : 1.f;
    for (int u = row_ptr[r] + lane; u < row_ptr[r + 1]; u += 32) {
      const float s = usym[u] / 2;
      const float th = utheta[u];
      float sn, cs;
      sincosf(two_pi_q * th, &sn, &cs);
      float mag;
      if (normalization) mag = (inv_sqrt_or_zero(deg[ucol[u]]) * s) * da;
      else mag = s;
      val_real[u] = scale_lmax(-(mag * cs), lambda_max);
      val_imag[u] = scale_lmax(mag * sn, lambda_max);
    }
  }
}

__global__ void k_magnetic_diag(const float* __restrict__ deg, int64_t N, int normalization,
                                float lambda_max, float* __restrict__ diag_real) {
  GRID_STRIDE(r, N) {
    const float d = normalization ? 1.0f : deg[r];
    diag_real[r] = scale_lmax(d, lambda_max) - 1.0f;
  }
}

// ---- generic / random-walk plans (32-bit keys = destination) ------------------------------
__global__ void k_keys_dst(const int64_t* __restrict__ dst, const int64_t* __restrict__ src,
                           int64_t E, int64_t n_dst, int64_t n_src, int drop_loops,
                           uint32_t sentinel, uint32_t* __restrict__ keys,
                           uint32_t* __restrict__ pos, int32_t* __restrict__ loop_pos,
                           Counters* cnt) {
  GRID_STRIDE(p, E) {
    const int64_t d = dst[p], s = src[p];
    uint32_t key = sentinel;
    if (d < 0 || d >= n_dst || s < 0 || s >= n_src) {
      cnt->err = 1;
    } else if (drop_loops && d == s) {
      atomicAdd(&cnt->n_loops, 1);
      atomicMax(&loop_pos[d], int32_t(p));  // the LAST self-loop of a node wins (index_put order)
    } else {
      key = uint32_t(d);
    }
    keys[p] = key;
    pos[p] = uint32_t(p);
  }
}

__global__ void k_fill_i32(int32_t* a, int64_t n, int32_t v) {
  GRID_STRIDE(i, n) a[i] = v;
}

__global__ void k_gather_sorted(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ pos,
                                const int64_t* __restrict__ src, const float* __restrict__ weight,
                                int64_t n, int32_t* __restrict__ rows, int32_t* __restrict__ col,
                                float* __restrict__ val) {
  GRID_STRIDE(k, n) {
    const uint32_t p = pos[k];
    rows[k] = int32_t(keys[k]);
    col[k] = int32_t(src[p]);
    if (val) val[k] = weight ? weight[p] : 1.f;
  }
}

// conv_norm_rw (conv_base.py:27-31): deg = rowsum(A_hat); w' = w / deg, inf -> 0.
__global__ void k_rw_scale(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ loop_pos,
                           const float* __restrict__ weight, float fill, int64_t N,
                           float* __restrict__ val, float* __restrict__ diag) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < N;
       r += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const int b = row_ptr[r], e = row_ptr[r + 1];
    float acc = 0.f;
    for (int k = b + lane; k < e; k += 32) acc += val[k];
    acc = warp_sum(acc);
    const int lp = loop_pos ? loop_pos[r] : -1;
    const float lw = lp >= 0 ? (weight ? weight[lp] : 1.f) : fill;
    float inv = 1.0f / (acc + lw);
    if (inv == INFINITY) inv = 0.f;
    for (int k = b + lane; k < e; k += 32) val[k] = inv * val[k];
    if (lane == 0) diag[r] = inv * lw;
  }
}

// gcn_norm (PyG; used by nn/directed/DGCNConv.py:75): deg = rowsum over the DESTINATION of
// A_hat, dis = deg^-1/2 (inf -> 0), w' = dis[src] * w * dis[dst].  Pass 1 leaves dis in `diag`.
__global__ void k_sym_deg(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ loop_pos,
                          const float* __restrict__ weight, float fill, int64_t N, const float* __restrict__ val,
                          float* __restrict__ dis) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < N;
       r += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    float acc = 0.f;
    for (int k = row_ptr[r] + lane; k < row_ptr[r + 1]; k += 32) acc += val[k];
    acc = warp_sum(acc);
    const int lp = loop_pos ? loop_pos[r] : -1;
    const float lw = loop_pos ? (lp >= 0 ? (weight ? weight[lp] : 1.f) : fill) : 0.f;
    if (lane == 0) dis[r] = inv_sqrt_or_zero(acc + lw);
  }
}
__global__ void k_sym_scale(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                            const float* __restrict__ dis, int64_t N, float* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  for (int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; r < N;
       r += (int64_t(gridDim.x) * blockDim.x) >> 5) {
    const float dr = dis[r];
    for (int k = row_ptr[r] + lane; k < row_ptr[r + 1]; k += 32) val[k] = (dis[col[k]] * val[k]) * dr;
  }
}
__global__ void k_sym_diag(const int32_t* __restrict__ loop_pos, const float* __restrict__ weight, float fill,
                           int64_t N, float* __restrict__ diag) {
  GRID_STRIDE(r, N) {
    const int lp = loop_pos ? loop_pos[r] : -1;
    const float lw = loop_pos ? (lp >= 0 ? (weight ? weight[lp] : 1.f) : fill) : 0.f;
    const float d = diag[r];            // dis[r] from k_sym_deg
    diag[r] = (d * lw) * d;
  }
}

// ---------------------------------------------------------------------------- workspace plan
struct Carver {
  char* base;
  size_t off = 0;
  explicit Carver(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

struct SymSpace {
  Counters* cnt;
  uint64_t *keys_a, *keys_b;
  uint32_t *slots_a, *slots_b;
  int32_t *head, *uidx, *urow;
  float *usym, *utheta, *uabs, *deg;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static size_t cub_temp_bound(int64_t M) {
  // DoubleBuffer radix sort needs only histogram + decoupled-lookback scratch: 256 counters per
  // tile of a few thousand keys (~0.3 B per key).  2 B per key + 8 MB is a safe bound and is
  // verified against CUB's own query at run time.
  return size_t(M) * 2 + (size_t(8) << 20);
}

static SymSpace carve_sym(void* ws, int64_t E, int64_t N) {
  const int64_t M = 2 * E;
  Carver c(ws);
  SymSpace s{};
  s.cnt = c.take<Counters>(1);
  s.keys_a = c.take<uint64_t>(M);
  s.keys_b = c.take<uint64_t>(M);
  s.slots_a = c.take<uint32_t>(M);
  s.slots_b = c.take<uint32_t>(M);
  s.head = c.take<int32_t>(M + 1);
  s.uidx = c.take<int32_t>(M + 1);
  s.urow = c.take<int32_t>(M + 1);
  s.usym = c.take<float>(M);
  s.utheta = c.take<float>(M);
  s.uabs = c.take<float>(M);
  s.deg = c.take<float>(N + 1);
  s.cub_bytes = cub_temp_bound(M);
  s.cub_temp = c.take<char>(s.cub_bytes);
  s.total = align_up(c.off);
  return s;
}

struct DstSpace {
  Counters* cnt;
  uint32_t *keys_a, *keys_b, *pos_a, *pos_b;
  int32_t *rows, *loop_pos;
  void* cub_temp;
  size_t cub_bytes;
  size_t total;
};

static DstSpace carve_dst(void* ws, int64_t E, int64_t N) {
  Carver c(ws);
  DstSpace s{};
  s.cnt = c.take<Counters>(1);
  s.keys_a = c.take<uint32_t>(E);
  s.keys_b = c.take<uint32_t>(E);
  s.pos_a = c.take<uint32_t>(E);
  s.pos_b = c.take<uint32_t>(E);
  s.rows = c.take<int32_t>(E + 1);
  s.loop_pos = c.take<int32_t>(N + 1);
  s.cub_bytes = cub_temp_bound(E);
  s.cub_temp = c.take<char>(s.cub_bytes);
  s.total = align_up(c.off);
  return s;
}

static int check_sizes(int64_t N, int64_t E) {
  PGSD_REQUIRE(N >= 0 && E >= 0, "plan: negative size");
  if (N >= (int64_t(1) << 31) - 1 || 2 * E >= (int64_t(1) << 31) - 1)
    return fail(PGSD_ERR_RANGE, "plan: N=%lld / 2E=%lld exceed the int32 plan format",
                (long long)N, (long long)(2 * E));
  return PGSD_OK;
}

static int read_counters(const Counters* dev, Counters* host, cudaStream_t st) {
  PGSD_CUDA(cudaMemcpyAsync(host, dev, sizeof(Counters), cudaMemcpyDeviceToHost, st));
  PGSD_CUDA(cudaStreamSynchronize(st));
  if (host->err) return fail(PGSD_ERR_INVALID, "plan: edge_index contains a node id outside [0, N)");
  return PGSD_OK;
}

// Shared tail of the two destination-keyed builders.
static int build_dst_sorted(const int64_t* dst, const int64_t* src, const float* w, int64_t E,
                            int64_t n_dst, int64_t n_src, int drop_loops, DstSpace& s,
                            int32_t* row_ptr, int32_t* col, float* val, Counters* host_cnt,
                            cudaStream_t st) {
  PGSD_CUDA(cudaMemsetAsync(s.cnt, 0, sizeof(Counters), st));
  if (drop_loops) {
    k_fill_i32<<<blocks_for(n_dst), TPB, 0, st>>>(s.loop_pos, n_dst, -1);
    PGSD_LAUNCH_CHECK("k_fill_i32");
  }
  const int bits = bit_length(uint64_t(n_dst)) + 1;
  const uint32_t sentinel = (bits >= 32) ? 0xffffffffu : ((1u << bits) - 1u);
  if (E > 0) {
    k_keys_dst<<<blocks_for(E), TPB, 0, st>>>(dst, src, E, n_dst, n_src, drop_loops, sentinel,
                                              s.keys_a, s.pos_a, s.loop_pos, s.cnt);
    PGSD_LAUNCH_CHECK("k_keys_dst");
    cub::DoubleBuffer<uint32_t> dk(s.keys_a, s.keys_b), dv(s.pos_a, s.pos_b);
    size_t need = 0;
    PGSD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, int(E), 0, bits, st));
    if (need > s.cub_bytes)
      return fail(PGSD_ERR_WORKSPACE, "plan: CUB scratch %zu > reserved %zu", need, s.cub_bytes);
    PGSD_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, need, dk, dv, int(E), 0, bits, st));
    int rc = read_counters(s.cnt, host_cnt, st);
    if (rc != PGSD_OK) return rc;
    host_cnt->nnz = int(E - host_cnt->n_loops);
    PGSD_CUDA(cudaMemcpyAsync(&s.cnt->nnz, &host_cnt->nnz, sizeof(int), cudaMemcpyHostToDevice, st));
    if (host_cnt->nnz > 0) {
      k_gather_sorted<<<blocks_for(host_cnt->nnz), TPB, 0, st>>>(dk.Current(), dv.Current(), src, w,
                                                                  host_cnt->nnz, s.rows, col, val);
      PGSD_LAUNCH_CHECK("k_gather_sorted");
    }
  } else {
    host_cnt->nnz = 0;
  }
  k_row_ptr<<<blocks_for(int64_t(host_cnt->nnz) + 1), TPB, 0, st>>>(s.rows, &s.cnt->nnz,
                                                                    host_cnt->nnz, n_dst, row_ptr);
  PGSD_LAUNCH_CHECK("k_row_ptr");
  return PGSD_OK;
}

}  // namespace pgsd

using namespace pgsd;

extern "C" int pgsd_plan_workspace_bytes(int64_t num_nodes, int64_t num_edges, size_t* bytes_host) {
  PGSD_REQUIRE(bytes_host != nullptr, "plan_workspace_bytes: null output");
  int rc = check_sizes(num_nodes, num_edges);
  if (rc != PGSD_OK) return rc;
  const size_t a = carve_sym(nullptr, num_edges, num_nodes).total;
  const size_t b = carve_dst(nullptr, num_edges, num_nodes).total;
  *bytes_host = (a > b ? a : b) + 256;
  return PGSD_OK;
}

extern "C" int pgsd_build_csr(const int64_t* edge_src, const int64_t* edge_dst,
                              const float* edge_weight, int64_t num_edges, int64_t num_dst,
                              int64_t num_src, int32_t* row_ptr, int32_t* col, float* val,
                              void* workspace, size_t workspace_bytes, pgsd_stream_t stream) {
  int rc = check_sizes(num_dst > num_src ? num_dst : num_src, num_edges);
  if (rc != PGSD_OK) return rc;
  PGSD_REQUIRE(row_ptr && (num_edges == 0 || (edge_src && edge_dst && col)), "build_csr: null pointer");
  PGSD_REQUIRE(workspace != nullptr, "build_csr: null workspace");
  DstSpace s = carve_dst(workspace, num_edges, num_dst);
  if (s.total > workspace_bytes)
    return fail(PGSD_ERR_WORKSPACE, "build_csr: workspace %zu < %zu", workspace_bytes, s.total);
  Counters hc{};
  return build_dst_sorted(edge_dst, edge_src, edge_weight, num_edges, num_dst, num_src, 0, s,
                          row_ptr, col, edge_weight ? val : nullptr, &hc,
                          static_cast<cudaStream_t>(stream));
}

extern "C" int pgsd_build_csr_rw_norm(const int64_t* edge_dst, const int64_t* edge_src,
                                      const float* edge_weight, int64_t num_edges,
                                      int64_t num_nodes, float fill_value, int add_self_loops,
                                      int32_t* row_ptr, int32_t* col, float* val, float* diag,
                                      int64_t* nnz_host,
                                      void* workspace, size_t workspace_bytes,
                                      pgsd_stream_t stream) {
  int rc = check_sizes(num_nodes, num_edges);
  if (rc != PGSD_OK) return rc;
  PGSD_REQUIRE(row_ptr && diag && nnz_host && (num_edges == 0 || (edge_src && edge_dst && col && val)),
               "build_csr_rw_norm: null pointer");
  PGSD_REQUIRE(workspace != nullptr, "build_csr_rw_norm: null workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DstSpace s = carve_dst(workspace, num_edges, num_nodes);
  if (s.total > workspace_bytes)
    return fail(PGSD_ERR_WORKSPACE, "build_csr_rw_norm: workspace %zu < %zu", workspace_bytes, s.total);
  Counters hc{};
  rc = build_dst_sorted(edge_dst, edge_src, edge_weight, num_edges, num_nodes, num_nodes,
                        add_self_loops ? 1 : 0, s, row_ptr, col, val, &hc, st);
  if (rc != PGSD_OK) return rc;
  *nnz_host = hc.nnz;
  if (num_nodes > 0) {
    k_rw_scale<<<blocks_for(num_nodes * 32), TPB, 0, st>>>(row_ptr, add_self_loops ? s.loop_pos : nullptr,
                                                           edge_weight, add_self_loops ? fill_value : 0.f,
                                                           num_nodes, val, diag);
    PGSD_LAUNCH_CHECK("k_rw_scale");
  }
  return PGSD_OK;
}

extern "C" int pgsd_build_csr_sym_norm(const int64_t* edge_dst, const int64_t* edge_src,
                                      const float* edge_weight, int64_t num_edges, int64_t num_nodes,
                                      float fill_value, int add_self_loops, int32_t* row_ptr, int32_t* col,
                                      float* val, float* diag, int64_t* nnz_host, void* workspace,
                                      size_t workspace_bytes, pgsd_stream_t stream) {
  int rc = check_sizes(num_nodes, num_edges);
  if (rc != PGSD_OK) return rc;
  PGSD_REQUIRE(row_ptr && diag && nnz_host && (num_edges == 0 || (edge_src && edge_dst && col && val)),
               "build_csr_sym_norm: null pointer");
  PGSD_REQUIRE(workspace != nullptr, "build_csr_sym_norm: null workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DstSpace s = carve_dst(workspace, num_edges, num_nodes);
  if (s.total > workspace_bytes)
    return fail(PGSD_ERR_WORKSPACE, "build_csr_sym_norm: workspace %zu < %zu", workspace_bytes, s.total);
  Counters hc{};
  rc = build_dst_sorted(edge_dst, edge_src, edge_weight, num_edges, num_nodes, num_nodes,
                        add_self_loops ? 1 : 0, s, row_ptr, col, val, &hc, st);
  if (rc != PGSD_OK) return rc;
  *nnz_host = hc.nnz;
  if (num_nodes > 0) {
    const int32_t* lp = add_self_loops ? s.loop_pos : nullptr;
    k_sym_deg<<<blocks_for(num_nodes * 32), TPB, 0, st>>>(row_ptr, lp, edge_weight, fill_value, num_nodes, val, diag);
    PGSD_LAUNCH_CHECK("k_sym_deg");
    k_sym_scale<<<blocks_for(num_nodes * 32), TPB, 0, st>>>(row_ptr, col, diag, num_nodes, val);
    PGSD_LAUNCH_CHECK("k_sym_scale");
    k_sym_diag<<<blocks_for(num_nodes), TPB, 0, st>>>(lp, edge_weight, fill_value, num_nodes, diag);
    PGSD_LAUNCH_CHECK("k_sym_diag");
  }
  return PGSD_OK;
}

static int build_magnetic_impl(
    const int64_t* edge_row, const int64_t* edge_col, const float* edge_weight, int64_t num_edges,
    int64_t num_nodes, double q, int normalization, float lambda_max, int signed_mode,
    int32_t* row_ptr, int32_t* col, float* val_real, float* val_imag, float* diag_real, float* theta,
    int64_t* nnz_host, void* workspace, size_t workspace_bytes, pgsd_stream_t stream,
    int64_t row_lo = 0, int64_t row_hi = -1, float* deg_out = nullptr) {
  // deg_out != nullptr: structure phase of the row-range build -- rows [row_lo, row_hi) only,
  // val_real / val_imag receive the coalesced A+A^T weight and Theta of every entry, deg_out the
  // row sums; the values follow in pgsd_build_magnetic_rows_finish once deg is known for all nodes.
  const bool structure_only = deg_out != nullptr;
  if (row_hi < 0) row_hi = num_nodes;
  int rc = check_sizes(num_nodes, num_edges);
  if (rc != PGSD_OK) return rc;
  PGSD_REQUIRE(row_ptr && (diag_real || structure_only) && nnz_host, "build_magnetic_laplacian: null pointer");
  PGSD_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= num_nodes, "build_magnetic_laplacian: bad row range");
  PGSD_REQUIRE(num_edges == 0 || (edge_row && edge_col && col && val_real && val_imag),
               "build_magnetic_laplacian: null pointer");
  PGSD_REQUIRE(signed_mode >= 0 && signed_mode <= 2, "build_magnetic_laplacian: bad signed_mode");
  PGSD_REQUIRE(workspace != nullptr, "build_magnetic_laplacian: null workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t E = num_edges, N = num_nodes, M = 2 * E, NL = row_hi - row_lo;
  SymSpace s = carve_sym(workspace, E, N);
  if (s.total > workspace_bytes)
    return fail(PGSD_ERR_WORKSPACE, "build_magnetic_laplacian: workspace %zu < %zu",
                workspace_bytes, s.total);
  PGSD_CUDA(cudaMemsetAsync(s.cnt, 0, sizeof(Counters), st));
  Counters hc{};
  if (M > 0) {
    const int bits = bit_length(uint64_t(NL > 0 ? NL : 1) * uint64_t(N)) + 1;
    const uint64_t sentinel = (uint64_t(1) << bits) - 1;
    k_keys_sym<<<blocks_for(M), TPB, 0, st>>>(edge_row, edge_col, E, N, row_lo, row_hi, sentinel, s.keys_a,
                                              s.slots_a, s.cnt);
    PGSD_LAUNCH_CHECK("k_keys_sym");
    cub::DoubleBuffer<uint64_t> dk(s.keys_a, s.keys_b);
    cub::DoubleBuffer<uint32_t> dv(s.slots_a, s.slots_b);
    size_t need = 0;
    PGSD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, int(M), 0, bits, st));
    if (need > s.cub_bytes)
      return fail(PGSD_ERR_WORKSPACE, "plan: CUB scratch %zu > reserved %zu", need, s.cub_bytes);
    PGSD_CUDA(cub::DeviceRadixSort::SortPairs(s.cub_temp, need, dk, dv, int(M), 0, bits, st));
    const uint64_t* keys = dk.Current();
    const uint32_t* slots = dv.Current();
    k_heads64<<<blocks_for(M), TPB, 0, st>>>(keys, M, sentinel, s.head);
    PGSD_LAUNCH_CHECK("k_heads64");
    need = 0;
    PGSD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, need, s.head, s.uidx, int(M), st));
    if (need > s.cub_bytes)
      return fail(PGSD_ERR_WORKSPACE, "plan: CUB scan scratch %zu > reserved %zu", need, s.cub_bytes);
    PGSD_CUDA(cub::DeviceScan::ExclusiveSum(s.cub_temp, need, s.head, s.uidx, int(M), st));
    k_reduce_dups<<<blocks_for(M), TPB, 0, st>>>(keys, slots, s.head, s.uidx, edge_weight, E, M, N,
                                                 s.urow, col, s.usym, s.utheta, s.uabs);
    PGSD_LAUNCH_CHECK("k_reduce_dups");
    k_set_nnz<<<1, 32, 0, st>>>(s.head, s.uidx, M, s.cnt);
    PGSD_LAUNCH_CHECK("k_set_nnz");
  }
  rc = read_counters(s.cnt, &hc, st);
  if (rc != PGSD_OK) return rc;
  *nnz_host = hc.nnz;
  k_row_ptr<<<blocks_for(int64_t(hc.nnz) + 1), TPB, 0, st>>>(s.urow, &s.cnt->nnz, hc.nnz, NL, row_ptr);
  PGSD_LAUNCH_CHECK("k_row_ptr");
  if (structure_only) {
    if (NL > 0) {
      k_degree<<<blocks_for(NL * 32), TPB, 0, st>>>(row_ptr, s.usym, s.uabs, signed_mode, NL, deg_out);
      PGSD_LAUNCH_CHECK("k_degree");
    }
    if (hc.nnz > 0) {
      PGSD_CUDA(cudaMemcpyAsync(val_real, s.usym, size_t(hc.nnz) * 4, cudaMemcpyDeviceToDevice, st));
      PGSD_CUDA(cudaMemcpyAsync(val_imag, s.utheta, size_t(hc.nnz) * 4, cudaMemcpyDeviceToDevice, st));
    }
    return PGSD_OK;
  }
  PGSD_REQUIRE(row_lo == 0 && row_hi == N, "build_magnetic_laplacian: a row range needs the two-phase build");
  if (N > 0) {
    k_degree<<<blocks_for(N * 32), TPB, 0, st>>>(row_ptr, s.usym, s.uabs, signed_mode, N, s.deg);
    PGSD_LAUNCH_CHECK("k_degree");
    k_magnetic_diag<<<blocks_for(N), TPB, 0, st>>>(s.deg, N, normalization, lambda_max, diag_real);
    PGSD_LAUNCH_CHECK("k_magnetic_diag");
  }
  if (hc.nnz > 0) {
    // 1j*2*pi*q is cast to complex64 before multiplying theta (get_magnetic_Laplacian.py:68):
    // the phase argument is the fp32 product fl(2*pi*q) * theta  (SURVEY Q11).
    const float two_pi_q = float(2.0 * 3.14159265358979323846 * q);
    k_magnetic_values<<<blocks_for(hc.nnz), TPB, 0, st>>>(s.urow, col, s.usym, s.utheta, s.deg,
                                                          &s.cnt->nnz, two_pi_q, normalization,
                                                          lambda_max, val_real, val_imag, theta);
    PGSD_LAUNCH_CHECK("k_magnetic_values");
  }
  return PGSD_OK;
}

extern "C" int pgsd_build_magnetic_laplacian(
    const int64_t* edge_row, const int64_t* edge_col, const float* edge_weight, int64_t num_edges,
    int64_t num_nodes, double q, int normalization, float lambda_max, int signed_mode,
    int32_t* row_ptr, int32_t* col, float* val_real, float* val_imag, float* diag_real,
    int64_t* nnz_host, void* workspace, size_t workspace_bytes, pgsd_stream_t stream) {
  return build_magnetic_impl(edge_row, edge_col, edge_weight, num_edges, num_nodes, q, normalization,
                             lambda_max, signed_mode, row_ptr, col, val_real, val_imag, diag_real,
                             nullptr, nnz_host, workspace, workspace_bytes, stream);
}

extern "C" int pgsd_build_magnetic_laplacian_theta(
    const int64_t* edge_row, const int64_t* edge_col, const float* edge_weight, int64_t num_edges,
    int64_t num_nodes, double q, int normalization, float lambda_max, int signed_mode,
    int32_t* row_ptr, int32_t* col, float* val_real, float* val_imag, float* diag_real, float* theta,
    int64_t* nnz_host, void* workspace, size_t workspace_bytes, pgsd_stream_t stream) {
  PGSD_REQUIRE(num_edges == 0 || theta, "build_magnetic_laplacian_theta: null theta");
  return build_magnetic_impl(edge_row, edge_col, edge_weight, num_edges, num_nodes, q, normalization,
                             lambda_max, signed_mode, row_ptr, col, val_real, val_imag, diag_real,
                             theta, nnz_host, workspace, workspace_bytes, stream);
}

extern "C" int pgsd_build_magnetic_rows_begin(
    const int64_t* edge_row, const int64_t* edge_col, const float* edge_weight, int64_t num_edges,
    int64_t num_nodes, int64_t row_lo, int64_t row_hi, int signed_mode, int32_t* row_ptr, int32_t* col,
    float* sym, float* theta, float* deg_local, int64_t* nnz_host, void* workspace, size_t workspace_bytes,
    pgsd_stream_t stream) {
  PGSD_REQUIRE(deg_local != nullptr, "build_magnetic_rows_begin: null deg_local");
  return build_magnetic_impl(edge_row, edge_col, edge_weight, num_edges, num_nodes, 0.0, 0, 2.0f, signed_mode,
                             row_ptr, col, sym, theta, nullptr, nullptr, nnz_host, workspace, workspace_bytes,
                             stream, row_lo, row_hi, deg_local);
}

extern "C" int pgsd_build_magnetic_rows_finish(
    const int32_t* row_ptr, const int32_t* col, const float* sym, const float* theta, const float* deg_all,
    int64_t num_nodes, int64_t row_lo, int64_t row_hi, double q, int normalization, float lambda_max,
    float* val_real, float* val_imag, float* diag_real, pgsd_stream_t stream) {
  PGSD_REQUIRE(0 <= row_lo && row_lo <= row_hi && row_hi <= num_nodes, "build_magnetic_rows_finish: bad row range");
  const int64_t NL = row_hi - row_lo;
  if (NL == 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr && deg_all && diag_real, "build_magnetic_rows_finish: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const float two_pi_q = float(2.0 * 3.14159265358979323846 * q);
  k_magnetic_values_rows<<<blocks_for(NL * 32), TPB, 0, st>>>(row_ptr, col, sym, theta, deg_all, NL, row_lo,
                                                             two_pi_q, normalization, lambda_max, val_real,
                                                             val_imag);
  PGSD_LAUNCH_CHECK("k_magnetic_values_rows");
  k_magnetic_diag<<<blocks_for(NL), TPB, 0, st>>>(deg_all + row_lo, NL, normalization, lambda_max, diag_real);
  PGSD_LAUNCH_CHECK("k_magnetic_diag");
  return PGSD_OK;
}
