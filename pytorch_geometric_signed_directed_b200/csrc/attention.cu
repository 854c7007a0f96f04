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
// One warp per destination row; up to two edge types, each with its own CSR plan.
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

__global__ void __launch_bounds__(256) edge_softmax_kernel(const AttnParams p) {
  const int lane = threadIdx.x & 31;
  const int64_t warps_total = int64_t(gridDim.x) * 8;
  for (int64_t row = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5); row < p.n_rows; row += warps_total) {
    int b[2] = {0, 0}, e[2] = {0, 0};
    float sd[2] = {0.f, 0.f};
    for (int t = 0; t < p.n_types; ++t) {
      b[t] = p.row_ptr[t][row], e[t] = p.row_ptr[t][row + 1];
      sd[t] = p.s_dst[t][row];
    }
    // pass 1: row maximum
    float m = -INFINITY;
    for (int t = 0; t < p.n_types; ++t)
      for (int k = b[t] + lane; k < e[t]; k += 32)
        m = fmaxf(m, attn_act(p.s_src[t][p.col[t][k]] + sd[t], p.act, p.slope));
    m = warp_max(m);
    // pass 2: exponent sums per type
    float sum[2] = {0.f, 0.f};
    for (int t = 0; t < p.n_types; ++t) {
      float acc = 0.f;
      for (int k = b[t] + lane; k < e[t]; k += 32)
        acc += expf(attn_act(p.s_src[t][p.col[t][k]] + sd[t], p.act, p.slope) - m);
      sum[t] = warp_add(acc);
    }
    const float inv = 1.0f / (sum[0] + sum[1] + 1e-16f);
    // mode B: per-entry alpha
    for (int t = 0; t < p.n_types; ++t)
      if (p.alpha_out[t] != nullptr)
        for (int k = b[t] + lane; k < e[t]; k += 32)
          p.alpha_out[t][k] = expf(attn_act(p.s_src[t][p.col[t][k]] + sd[t], p.act, p.slope) - m) * inv;
    // mode A: y[i] = xd0[i] * P + xd1[i] * N
    if (p.y != nullptr) {
      const bool any = (e[0] - b[0]) + (e[1] - b[1]) > 0;
      const float w0 = any ? sum[0] * inv : 0.f, w1 = any ? sum[1] * inv : 0.f;
      for (int f = lane; f < p.feat; f += 32) {
        float v = w0 * p.xd[0][row * p.ldxd[0] + f];
        if (p.n_types == 2) v = fmaf(w1, p.xd[1][row * p.ldxd[1] + f], v);
        p.y[row * p.ldy + f] = v;
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
  int64_t grid = ceil_div<int64_t>(a->n_rows, 8);
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  edge_softmax_kernel<<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  PGSD_LAUNCH_CHECK("edge_softmax_kernel");
  return PGSD_OK;
}
