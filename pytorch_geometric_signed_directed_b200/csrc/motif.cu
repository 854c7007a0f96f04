// Signed-triangle motif counts for SDGNN / SiGAT: for every signed edge (u, v) the 16 numbers
//   d[4*a + b] = | N_{lu(a,b)}(u)  intersect  N_{lv(a,b)}(v) |
// over the four neighbour lists pos_out, pos_in, neg_out, neg_in (as SETS), in the order of the reference's
// get_features() / get_tri_features() tuple (nn/signed/SDGNN.py:153-195, nn/signed/SiGAT.py:93-134).  The
// reference walks Python sets edge by edge on the CPU; here one warp owns an edge, the lanes take the elements
// of the shorter sorted list and binary-search the longer one.  Integer work: results are exact.
#include "common.cuh"

namespace pgsd {

struct MotifParams {
  const int32_t* row_ptr[4];   // 0 pos_out, 1 pos_in, 2 neg_out, 3 neg_in: sorted, duplicate-free rows
  const int32_t* col[4];
  const int64_t* eu;
  const int64_t* ev;
  int64_t n_edges, n_nodes;
  int32_t* out;                // [n_edges, 16]
};

// (list of u, list of v) for d1_1..d1_4, d2_1..d2_4, d3_1..d3_4, d4_1..d4_4
__constant__ int8_t MOTIF_LU[16] = {0, 0, 2, 2, 0, 0, 2, 2, 1, 1, 3, 3, 1, 1, 3, 3};
__constant__ int8_t MOTIF_LV[16] = {1, 3, 1, 3, 0, 2, 0, 2, 0, 2, 0, 2, 1, 3, 1, 3};

__device__ __forceinline__ bool sorted_contains(const int32_t* __restrict__ a, int n, int32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    const int32_t v = __ldg(a + mid);
    if (v < key) lo = mid + 1; else hi = mid;
  }
  return lo < n && __ldg(a + lo) == key;
}

__global__ void __launch_bounds__(256) signed_triangle_counts_kernel(const MotifParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = (int64_t(gridDim.x) * blockDim.x) >> 5;
  for (int64_t e = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5; e < p.n_edges; e += warps) {
    const int64_t u = p.eu[e], v = p.ev[e];
    const bool valid = u >= 0 && u < p.n_nodes && v >= 0 && v < p.n_nodes;
#pragma unroll 1
    for (int t = 0; t < 16; ++t) {
      int cnt = 0;
      if (valid) {
        const int lu = MOTIF_LU[t], lv = MOTIF_LV[t];
        const int as = __ldg(p.row_ptr[lu] + u), an = __ldg(p.row_ptr[lu] + u + 1) - as;
        const int bs = __ldg(p.row_ptr[lv] + v), bn = __ldg(p.row_ptr[lv] + v + 1) - bs;
        const int32_t* a = p.col[lu] + as;
        const int32_t* b = p.col[lv] + bs;
        const int32_t* sh = an <= bn ? a : b;      // lanes walk the shorter list ...
        const int32_t* lg = an <= bn ? b : a;      // ... and search the longer one
        const int sn = an <= bn ? an : bn, ln = an <= bn ? bn : an;
        if (ln > 0)
          for (int i = lane; i < sn; i += 32) cnt += sorted_contains(lg, ln, __ldg(sh + i)) ? 1 : 0;
      }
      cnt = __reduce_add_sync(0xffffffffu, cnt);
      if (lane == 0) p.out[e * 16 + t] = cnt;
    }
  }
}

}  // namespace pgsd

using namespace pgsd;

extern "C" int pgsd_signed_triangle_counts(const int32_t* const* row_ptr4, const int32_t* const* col4,
                                           const int64_t* edge_u, const int64_t* edge_v, int64_t n_edges,
                                           int64_t n_nodes, int32_t* counts, pgsd_stream_t stream) {
  PGSD_REQUIRE(n_edges >= 0 && n_nodes >= 0, "signed_triangle_counts: negative size");
  if (n_edges == 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr4 && col4 && edge_u && edge_v && counts, "signed_triangle_counts: null pointer");
  MotifParams p{};
  for (int k = 0; k < 4; ++k) {
    PGSD_REQUIRE(row_ptr4[k] != nullptr, "signed_triangle_counts: row_ptr[%d] is null", k);
    p.row_ptr[k] = row_ptr4[k];
    p.col[k] = col4[k];                            // may be null for an empty list
  }
  p.eu = edge_u, p.ev = edge_v, p.n_edges = n_edges, p.n_nodes = n_nodes, p.out = counts;
  int64_t grid = ceil_div<int64_t>(n_edges, 8);
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  signed_triangle_counts_kernel<<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  PGSD_LAUNCH_CHECK("signed_triangle_counts_kernel");
  return PGSD_OK;
}
