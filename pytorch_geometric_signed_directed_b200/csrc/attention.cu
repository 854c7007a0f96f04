// Segment-softmax attention over CSR-by-destination plans (SNEAConv; GATConv-style layers).
//
// Reference: nn/signed/SNEAConv.py:135-146 (message) + PyG utils.softmax:
//   t_e     = act( s_src[p_e][j_e] + s_dst[p_e][i] )            per edge e = (j -> i) of type p_e
//   alpha_e = exp(t_e - max_i) / (sum_{e' -> i} exp(t_e' - max_i) + 1e-16)
// where the per-node scalars are the two halves of the reference's Linear(2*out -> 1) applied
// to [x_j || x_i] (s_src = X a_j, s_dst = X a_i + c), so the per-edge work is two SCALAR
// gathers instead of two F-wide gathers + a [nnz, 2F] temporary.
//   mode A (SNEAConv, reference quirk Q7: the message is the TARGET's feature times alpha):
//       y[i] = xd0[i] * sum_{e in type 0} alpha_e + xd1[i] * sum_{e in type 1} alpha_e
//   mode B (GATConv-style): alpha_e is written per stored entry; the weighted aggregation
//       sum alpha_e x_j then runs through pgsd_spmm_csr with val = alpha.
// One lane group per destination row; up to two edge types, each with its own CSR plan.
#include "common.cuh"

namespace pgsd {

struct AttnParams {
  int64_t n_rows;
  int32_t feat, n_types, act;
  float slope;
  const int32_t* row_ptr[2];
  const int32_t* col[2];
  const float* s_src[2];
  const float* s_dst[2];
  const float* xd[2];
  int64_t ldxd[2];
  float* y;
  int64_t ldy;
  float* alpha_out[2];
};

__device__ __forceinline__ float attn_act(float v, int act, float slope) {
  if (act == 0) return tanhf(v);
  return v > 0.f ? v : slope * v;   // leaky_relu
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_add(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One LPR-lane group per destination row (rows have ~10-20 entries: a whole warp per row leaves most lanes
// idle), ONE pass over the entries for the softmax statistics: every lane keeps a running maximum and the
// exponent sums relative to it (rescaled when the maximum moves), the group merges the (max, sums) triples
// with shuffles.  Mode A needs nothing else; mode B revisits the entries once to write alpha.  The first
// version of this kernel used a warp per row and re-gathered / re-evaluated every entry in three passes
// (0.75 ms per 20M entries; this one: see profiles/README.md).
template <int LPR>
__global__ void __launch_bounds__(256) edge_softmax_kernel(const AttnParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
  const int64_t warp_id = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t stride = int64_t(gridDim.x) * 8 * G;
  const bool vec = (p.feat % 4 == 0) && p.y != nullptr && (reinterpret_cast<uintptr_t>(p.y) % 16 == 0) &&
                   (p.ldy % 4 == 0) && (reinterpret_cast<uintptr_t>(p.xd[0]) % 16 == 0) && (p.ldxd[0] % 4 == 0) &&
                   (p.n_types == 1 || ((reinterpret_cast<uintptr_t>(p.xd[1]) % 16 == 0) && (p.ldxd[1] % 4 == 0)));
  for (int64_t base = warp_id * G; base < p.n_rows; base += stride) {      // warp-uniform trip count
    const int64_t row = base + g;
    const bool live = row < p.n_rows;
    int b[2] = {0, 0}, e[2] = {0, 0};
    float sd[2] = {0.f, 0.f};
    if (live)
      for (int t = 0; t < p.n_types; ++t) {
        b[t] = __ldg(p.row_ptr[t] + row), e[t] = __ldg(p.row_ptr[t] + row + 1);
        sd[t] = __ldg(p.s_dst[t] + row);
      }
    float m = -INFINITY, sum[2] = {0.f, 0.f};
    for (int t = 0; t < p.n_types; ++t)
      for (int k = b[t] + l; k < e[t]; k += LPR) {
        const float v = attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope);
        if (v > m) {
          const float sc = expf(m - v);          // exp(-inf) = 0 on the first entry
          sum[0] *= sc, sum[1] *= sc, m = v;
        }
        sum[t] += expf(v - m);
      }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      const float m2 = __shfl_xor_sync(FULL, m, o, LPR);
      const float a0 = __shfl_xor_sync(FULL, sum[0], o, LPR), a1 = __shfl_xor_sync(FULL, sum[1], o, LPR);
      const float mn = fmaxf(m, m2);
      const float f1 = (m == -INFINITY) ? 0.f : expf(m - mn), f2 = (m2 == -INFINITY) ? 0.f : expf(m2 - mn);
      sum[0] = sum[0] * f1 + a0 * f2, sum[1] = sum[1] * f1 + a1 * f2, m = mn;
    }
    const float inv = 1.0f / (sum[0] + sum[1] + 1e-16f);
    // mode B: per-entry alpha
    for (int t = 0; t < p.n_types; ++t)
      if (p.alpha_out[t] != nullptr)
        for (int k = b[t] + l; k < e[t]; k += LPR)
          p.alpha_out[t][k] =
              expf(attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope) - m) * inv;
    // mode A: y[i] = xd0[i] * P + xd1[i] * N
    if (p.y != nullptr && live) {
      const float w0 = sum[0] * inv, w1 = sum[1] * inv;       // an isolated row: 0 * 1e16 = 0
      if (vec) {
        for (int f = l * 4; f < p.feat; f += LPR * 4) {
          const float4 x0 = __ldg(reinterpret_cast<const float4*>(p.xd[0] + row * p.ldxd[0] + f));
          float4 v = make_float4(w0 * x0.x, w0 * x0.y, w0 * x0.z, w0 * x0.w);
          if (p.n_types == 2) {
            const float4 x1 = __ldg(reinterpret_cast<const float4*>(p.xd[1] + row * p.ldxd[1] + f));
            v.x = fmaf(w1, x1.x, v.x), v.y = fmaf(w1, x1.y, v.y), v.z = fmaf(w1, x1.z, v.z), v.w = fmaf(w1, x1.w, v.w);
          }
          *reinterpret_cast<float4*>(p.y + row * p.ldy + f) = v;
        }
      } else {
        for (int f = l; f < p.feat; f += LPR) {
          float v = w0 * p.xd[0][row * p.ldxd[0] + f];
          if (p.n_types == 2) v = fmaf(w1, p.xd[1][row * p.ldxd[1] + f], v);
          p.y[row * p.ldy + f] = v;
        }
      }
    }
  }
}

}  // namespace pgsd

using namespace pgsd;

extern "C" int pgsd_edge_softmax(const pgsd_attn_args* a, pgsd_stream_t stream) {
  PGSD_REQUIRE(a != nullptr, "edge_softmax: args is null");
  PGSD_REQUIRE(a->n_types == 1 || a->n_types == 2, "edge_softmax: n_types must be 1 or 2");
  PGSD_REQUIRE(a->act == 0 || a->act == 1, "edge_softmax: act must be 0 (tanh) or 1 (leaky_relu)");
  PGSD_REQUIRE(a->n_rows >= 0 && a->feat >= 0, "edge_softmax: negative size");
  if (a->n_rows == 0) return PGSD_OK;
  AttnParams p{};
  p.n_rows = a->n_rows, p.feat = a->feat, p.n_types = a->n_types, p.act = a->act, p.slope = a->slope;
  for (int t = 0; t < a->n_types; ++t) {
    PGSD_REQUIRE(a->row_ptr[t] && a->s_src[t] && a->s_dst[t], "edge_softmax: null pointer (type %d)", t);
    p.row_ptr[t] = a->row_ptr[t], p.col[t] = a->col[t];
    p.s_src[t] = a->s_src[t], p.s_dst[t] = a->s_dst[t];
    p.alpha_out[t] = a->alpha_out[t];
    if (a->y) {
      PGSD_REQUIRE(a->xd[t] != nullptr, "edge_softmax: xd[%d] is null", t);
      p.xd[t] = a->xd[t], p.ldxd[t] = a->ldxd[t];
    }
  }
  p.y = a->y, p.ldy = a->ldy;
  constexpr int LPR = 8;                                    // 4 rows per warp
  int64_t grid = ceil_div<int64_t>(a->n_rows, 8 * (32 / LPR));
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  edge_softmax_kernel<LPR><<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  PGSD_LAUNCH_CHECK("edge_softmax_kernel");
  return PGSD_OK;
}
