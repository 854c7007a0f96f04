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
// One LPR-lane group per destination row (rows have ~10-20 entries: a whole warp per row leaves most lanes
// idle).  Every entry is gathered and activated ONCE: the first CACHE entries per lane and type stay in registers
// between the maximum pass and the sum pass (longer rows re-evaluate the rest).  The first version of this kernel
// used a warp per row and re-gathered / re-evaluated every entry in three passes (0.75 ms per 20M entries); a
// grouped version with online rescaling was issue-bound on its expf calls (0.36-0.41 ms, ncu session 41).
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
    // pass 1: activations of this lane's entries (the first CACHE per type stay in registers), group maximum.
    // No exponentials here: ncu showed the first grouped version issue-bound (68 % issue-active at 19 % DRAM)
    // on the expf calls of its online rescaling and of the shuffle merge, not on the gathers.
    constexpr int CACHE = 2;
    float cv[2][CACHE];
    float m = -INFINITY;
#pragma unroll
    for (int t = 0; t < 2; ++t) {                       // fixed bound: cv[][] must stay in registers
      if (t >= p.n_types) continue;
#pragma unroll
      for (int i = 0; i < CACHE; ++i) {
        const int k = b[t] + l + i * LPR;
        cv[t][i] = -INFINITY;
        if (k < e[t]) cv[t][i] = attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope);
        m = fmaxf(m, cv[t][i]);
      }
      for (int k = b[t] + l + CACHE * LPR; k < e[t]; k += LPR)            // long rows: not cached
        m = fmaxf(m, attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope));
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o, LPR));
    // pass 2: exponent sums per type (one expf per entry), group sums
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 2; ++t) {
      if (t >= p.n_types) continue;
#pragma unroll
      for (int i = 0; i < CACHE; ++i)
        if (cv[t][i] != -INFINITY) sum[t] += expf(cv[t][i] - m);
      for (int k = b[t] + l + CACHE * LPR; k < e[t]; k += LPR)
        sum[t] += expf(attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope) - m);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) {
      sum[0] += __shfl_xor_sync(FULL, sum[0], o, LPR);
      sum[1] += __shfl_xor_sync(FULL, sum[1], o, LPR);
    }
    const float inv = 1.0f / (sum[0] + sum[1] + 1e-16f);
    // mode B: per-entry alpha
#pragma unroll
    for (int t = 0; t < 2; ++t)
      if (t < p.n_types && p.alpha_out[t] != nullptr) {
#pragma unroll
        for (int i = 0; i < CACHE; ++i) {
          const int k = b[t] + l + i * LPR;
          if (k < e[t]) p.alpha_out[t][k] = expf(cv[t][i] - m) * inv;
        }
        for (int k = b[t] + l + CACHE * LPR; k < e[t]; k += LPR)
          p.alpha_out[t][k] =
              expf(attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope) - m) * inv;
      }
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

// ---- GATConv aggregation with the softmax inside (SDRLayer / SiGAT) -----------------------
//   y[i] = sum_e alpha_e h[j_e] (+ bias) (+ beta z[i]),   alpha_e = exp(t_e - max_i) / (sum_e' exp(t_e' - max_i) + 1e-16),
//   t_e = leaky_relu(s_src[j_e] + s_dst[i])
// One LPR-lane group per destination row, like the aggregation kernels of spmm.cu, but the per-entry weight is
// computed on the fly from two scalar gathers and normalised at the END of the row (running maximum, rescaled
// partial sums): the separate softmax launch, the alpha array and its re-read disappear.
struct GatParams {
  int64_t n_rows;
  int32_t feat, lpr_active;
  float slope, beta;
  const int32_t* row_ptr;
  const int32_t* col;
  const float* s_src;
  const float* s_dst;
  const float* h;
  int64_t ldh;
  const float* bias;
  const float* z;
  int64_t ldz;
  float* y;
  int64_t ldy;
};

// Software pipeline as in spmm_groups_kernel: while the feature rows of the current batch are being gathered,
// the column indices and then the scores of the NEXT batch (same row, or the first batch of the group's next
// row, whose row_ptr pair / s_dst were fetched one row ahead) are already in flight -- without it the chain
// row_ptr -> col -> s_src -> exp -> gathers is exposed once per row and the kernel is no faster than the
// softmax + aggregation pair it replaces (measured, session 39).
template <int LPR>
__global__ void __launch_bounds__(256) gat_aggregate_kernel(const GatParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = 32 / LPR, U = 4;
  const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
  const bool lane_active = l < p.lpr_active;
  const int64_t stride = int64_t(gridDim.x) * 8 * G;
  const float* hb = p.h + l * 4;

  auto load_row = [&](int64_t r, int& b, int& e, float& sd) {
    b = e = 0, sd = 0.f;
    if (r < p.n_rows) b = __ldg(p.row_ptr + r), e = __ldg(p.row_ptr + r + 1), sd = __ldg(p.s_dst + r);
  };
  auto score = [&](int c, float sd) {
    const float v = __ldg(p.s_src + c) + sd;
    return v > 0.f ? v : p.slope * v;
  };

  int64_t row = (int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5)) * G + g;
  int b, e, nb, ne;
  float sd, nsd;
  load_row(row, b, e, sd);
  int64_t nrow = row + stride;
  load_row(nrow, nb, ne, nsd);
  int kb = b;
  int c = 0;
  float t = -INFINITY;
  if (kb + l < e) c = __ldg(p.col + kb + l), t = score(c, sd);

  float m = -INFINITY, lsum = 0.f;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);

  while (__any_sync(FULL, row < p.n_rows)) {
    const int cnt = min(LPR, e - kb);                       // <= 0: empty / finished row
    const bool more = kb + LPR < e;
    // next batch: issue its column-index load now
    const int nk = (more ? kb + LPR : nb) + l;
    const int nend = more ? e : ne;
    const float nsdv = more ? sd : nsd;
    const bool nvalid = nk < nend;
    const int nc = nvalid ? __ldg(p.col + nk) : 0;

    // softmax statistics of the current batch
    float bm = t;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) bm = fmaxf(bm, __shfl_xor_sync(FULL, bm, o, LPR));
    const float mn = fmaxf(m, bm);
    const float sc = (m == -INFINITY) ? 0.f : expf(m - mn);
    acc.x *= sc, acc.y *= sc, acc.z *= sc, acc.w *= sc, lsum *= sc;
    m = mn;
    const float pw = (l < cnt) ? expf(t - mn) : 0.f;
    float ps = pw;
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) ps += __shfl_xor_sync(FULL, ps, o, LPR);
    lsum += ps;

    float nt = -INFINITY;
    for (int j = 0; __any_sync(FULL, j < cnt); j += U) {
      float4 d[U];
      float w[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int idx = j + u;
        const int cc = __shfl_sync(FULL, c, idx & (LPR - 1), LPR);
        const float ww = __shfl_sync(FULL, pw, idx & (LPR - 1), LPR);
        const bool ok = idx < cnt && lane_active;
        w[u] = ok ? ww : 0.f;
        if (ok) d[u] = __ldg(reinterpret_cast<const float4*>(hb + int64_t(cc) * p.ldh));
        else d[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        acc.x = fmaf(w[u], d[u].x, acc.x), acc.y = fmaf(w[u], d[u].y, acc.y);
        acc.z = fmaf(w[u], d[u].z, acc.z), acc.w = fmaf(w[u], d[u].w, acc.w);
      }
    }
    // The next batch's scores go out LAST: they depend on nc, and a dependent load in the middle of the gather
    // sequence stalls the in-order warp before the remaining gathers are issued (2.0 ms per launch instead of
    // 1.46, session 39); here nc has arrived behind the gathers and only the score latency is exposed.
    if (nvalid) nt = score(nc, nsdv);

    if (more) {
      kb += LPR;
    } else {
      if (row < p.n_rows && lane_active) {
        const float inv = 1.0f / (lsum + 1e-16f);
        float4 o = make_float4(acc.x * inv, acc.y * inv, acc.z * inv, acc.w * inv);
        if (p.bias) {
          const float4 bs = __ldg(reinterpret_cast<const float4*>(p.bias + l * 4));
          o.x += bs.x, o.y += bs.y, o.z += bs.z, o.w += bs.w;
        }
        if (p.z) {
          const float4 zz = *reinterpret_cast<const float4*>(p.z + row * p.ldz + l * 4);
          o.x = fmaf(p.beta, zz.x, o.x), o.y = fmaf(p.beta, zz.y, o.y), o.z = fmaf(p.beta, zz.z, o.z),
          o.w = fmaf(p.beta, zz.w, o.w);
        }
        *reinterpret_cast<float4*>(p.y + row * p.ldy + l * 4) = o;
      }
      m = -INFINITY, lsum = 0.f, acc = make_float4(0.f, 0.f, 0.f, 0.f);
      row = nrow, b = nb, e = ne, sd = nsd, kb = nb;
      nrow += stride;
      load_row(nrow, nb, ne, nsd);
    }
    c = nc, t = nt;
    __syncwarp();
  }
}

// ---- backward of the segment softmax (training through SNEAConv / GATConv) -----------------------------------
// Reference: autograd through nn/signed/SNEAConv.py:135-146 and PyG GATConv as used by nn/signed/SDGNN.py:35-64.
// Per row i the kernel re-evaluates the softmax (max, sums) and then, with dalpha_e = dL/dalpha_e,
//   D      = sum_e alpha_e dalpha_e
//   dt_e   = alpha_e (dalpha_e - D)                       (softmax backward; the 1e-16 in the denominator is kept)
//   dpre_e = dt_e * act'(pre_e)                           (tanh: 1 - t^2;  leaky_relu: 1 or slope)
//   g_s_src[p][col_e] += dpre_e  (atomics: a source is shared by many rows);   g_s_dst[p][i] = sum_{e of type p} dpre_e
// dalpha_e comes either per entry (GAT-style: the SDDMM below) or as one coefficient per row and type
// (SNEAConv: y[i] = xd0[i] S0 + xd1[i] S1 => dalpha_e = <gy[i], xd_type(e)[i]>).  The per-row type sums S_p are written
// out as well: they are the SNEAConv gradient w.r.t. xd (g_xd_p = gy * S_p).
struct AttnBwdParams {
  int64_t n_rows;
  int32_t n_types, act;
  float slope;
  const int32_t* row_ptr[2];
  const int32_t* col[2];
  const float* s_src[2];
  const float* s_dst[2];
  const float* dalpha[2];     // per entry, or NULL
  const float* row_coef[2];   // per row, used when dalpha is NULL
  float* g_s_src[2];
  float* g_s_dst[2];
  float* type_sum[2];         // optional S_p per row
};

template <int LPR>
__global__ void __launch_bounds__(256) edge_softmax_bwd_kernel(const AttnBwdParams p) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int G = 32 / LPR;
  const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
  const int64_t warp_id = int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5);
  const int64_t stride = int64_t(gridDim.x) * 8 * G;
  auto gsum = [&](float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, LPR);
    return v;
  };
  for (int64_t base = warp_id * G; base < p.n_rows; base += stride) {
    const int64_t row = base + g;
    const bool live = row < p.n_rows;
    int b[2] = {0, 0}, e[2] = {0, 0};
    float sd[2] = {0.f, 0.f}, rc[2] = {0.f, 0.f};
    if (live)
      for (int t = 0; t < p.n_types; ++t) {
        b[t] = __ldg(p.row_ptr[t] + row), e[t] = __ldg(p.row_ptr[t] + row + 1);
        sd[t] = __ldg(p.s_dst[t] + row);
        if (p.dalpha[t] == nullptr && p.row_coef[t] != nullptr) rc[t] = __ldg(p.row_coef[t] + row);
      }
    float m = -INFINITY;
    for (int t = 0; t < p.n_types; ++t)
      for (int k = b[t] + l; k < e[t]; k += LPR)
        m = fmaxf(m, attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope));
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o, LPR));
    float sum[2] = {0.f, 0.f}, dot = 0.f;
    for (int t = 0; t < p.n_types; ++t)
      for (int k = b[t] + l; k < e[t]; k += LPR) {
        const float ex = expf(attn_act(__ldg(p.s_src[t] + __ldg(p.col[t] + k)) + sd[t], p.act, p.slope) - m);
        sum[t] += ex;
        dot += ex * (p.dalpha[t] ? __ldg(p.dalpha[t] + k) : rc[t]);
      }
    sum[0] = gsum(sum[0]), sum[1] = gsum(sum[1]), dot = gsum(dot);
    const float inv = 1.0f / (sum[0] + sum[1] + 1e-16f);
    const float D = dot * inv;
    float gd[2] = {0.f, 0.f};
    for (int t = 0; t < p.n_types; ++t)
      for (int k = b[t] + l; k < e[t]; k += LPR) {
        const int c = __ldg(p.col[t] + k);
        const float pre = __ldg(p.s_src[t] + c) + sd[t];
        const float a = attn_act(pre, p.act, p.slope);
        const float alpha = expf(a - m) * inv;
        const float da = p.dalpha[t] ? __ldg(p.dalpha[t] + k) : rc[t];
        const float dact = p.act == 0 ? (1.f - a * a) : (pre > 0.f ? 1.f : p.slope);
        const float dpre = alpha * (da - D) * dact;
        gd[t] += dpre;
        atomicAdd(p.g_s_src[t] + c, dpre);
      }
    gd[0] = gsum(gd[0]), gd[1] = gsum(gd[1]);
    if (live && l == 0)
      for (int t = 0; t < p.n_types; ++t) {
        p.g_s_dst[t][row] = gd[t];
        if (p.type_sum[t]) p.type_sum[t][row] = sum[t] * inv;
      }
  }
}

// SDDMM of the GAT-style backward: out[k] = <gy[row(k)], h[col[k]]> for every stored entry k (= dL/dalpha_k of
// y[i] = sum_k alpha_k h[col_k]).  One 16-lane group per row keeps its slice of gy[row] in registers and gathers the
// source rows like the aggregation kernels.  feat % 4 == 0, feat <= 256, 16-byte aligned rows.
__global__ void __launch_bounds__(256) sddmm_rows_kernel(const int32_t* __restrict__ row_ptr, const int32_t* __restrict__ col,
                                                         const float* __restrict__ gy, int64_t ldg,
                                                         const float* __restrict__ h, int64_t ldh, int64_t n_rows,
                                                         int feat, float* __restrict__ out) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int LPR = 16, G = 2, MAXV = 4;
  const int lane = threadIdx.x & 31, g = lane / LPR, l = lane % LPR;
  const int64_t stride = int64_t(gridDim.x) * 8 * G;
  for (int64_t base = (int64_t(blockIdx.x) * 8 + (threadIdx.x >> 5)) * G; base < n_rows; base += stride) {
    const int64_t row = base + g;
    const bool live = row < n_rows;
    int b = 0, e = 0;
    if (live) b = __ldg(row_ptr + row), e = __ldg(row_ptr + row + 1);
    float4 gr[MAXV];
#pragma unroll
    for (int v = 0; v < MAXV; ++v) {
      const int f = (v * LPR + l) * 4;
      gr[v] = (live && f < feat) ? __ldg(reinterpret_cast<const float4*>(gy + row * ldg + f)) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    int len = e - b;
    len = max(len, __shfl_xor_sync(FULL, len, LPR));        // both groups of the warp iterate together
    for (int k = 0; k < len; ++k) {
      float acc = 0.f;
      if (b + k < e) {
        const float* hr = h + int64_t(__ldg(col + b + k)) * ldh;
#pragma unroll
        for (int v = 0; v < MAXV; ++v) {
          const int f = (v * LPR + l) * 4;
          if (f < feat) {
            const float4 x = __ldg(reinterpret_cast<const float4*>(hr + f));
            acc = fmaf(gr[v].x, x.x, fmaf(gr[v].y, x.y, fmaf(gr[v].z, x.z, fmaf(gr[v].w, x.w, acc))));
          }
        }
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(FULL, acc, o, LPR);
      if (l == 0 && b + k < e) out[b + k] = acc;
    }
  }
}

}  // namespace pgsd

using namespace pgsd;

extern "C" size_t pgsd_sizeof_attn_bwd_args(void) { return sizeof(pgsd_attn_bwd_args); }

extern "C" int pgsd_edge_softmax_backward(const pgsd_attn_bwd_args* a, pgsd_stream_t stream) {
  PGSD_REQUIRE(a != nullptr, "edge_softmax_backward: args is null");
  PGSD_REQUIRE(a->n_types == 1 || a->n_types == 2, "edge_softmax_backward: n_types must be 1 or 2");
  PGSD_REQUIRE(a->act == 0 || a->act == 1, "edge_softmax_backward: act must be 0 (tanh) or 1 (leaky_relu)");
  if (a->n_rows <= 0) return PGSD_OK;
  AttnBwdParams p{};
  p.n_rows = a->n_rows, p.n_types = a->n_types, p.act = a->act, p.slope = a->slope;
  for (int t = 0; t < a->n_types; ++t) {
    PGSD_REQUIRE(a->row_ptr[t] && a->s_src[t] && a->s_dst[t] && a->g_s_src[t] && a->g_s_dst[t],
                 "edge_softmax_backward: null pointer (type %d)", t);
    p.row_ptr[t] = a->row_ptr[t], p.col[t] = a->col[t], p.s_src[t] = a->s_src[t], p.s_dst[t] = a->s_dst[t];
    p.dalpha[t] = a->dalpha[t], p.row_coef[t] = a->row_coef[t];
    p.g_s_src[t] = a->g_s_src[t], p.g_s_dst[t] = a->g_s_dst[t], p.type_sum[t] = a->type_sum[t];
  }
  constexpr int LPR = 8;
  int64_t grid = ceil_div<int64_t>(a->n_rows, 8 * (32 / LPR));
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  edge_softmax_bwd_kernel<LPR><<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  PGSD_LAUNCH_CHECK("edge_softmax_bwd_kernel");
  return PGSD_OK;
}

extern "C" int pgsd_sddmm_rows(const int32_t* row_ptr, const int32_t* col, const float* gy, int64_t ldg, const float* h,
                               int64_t ldh, int64_t n_rows, int32_t feat, float* out, pgsd_stream_t stream) {
  if (n_rows <= 0 || feat <= 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr && gy && h && out, "sddmm_rows: null pointer");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  PGSD_REQUIRE(feat % 4 == 0 && feat <= 256 && al16(gy) && al16(h) && ldg % 4 == 0 && ldh % 4 == 0,
               "sddmm_rows: feat must be a multiple of 4 (<= 256) with 16-byte aligned rows");
  int64_t grid = ceil_div<int64_t>(n_rows, 16);
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  sddmm_rows_kernel<<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(row_ptr, col, gy, ldg, h, ldh, n_rows,
                                                                                 feat, out);
  PGSD_LAUNCH_CHECK("sddmm_rows_kernel");
  return PGSD_OK;
}


extern "C" int pgsd_gat_aggregate(const int32_t* row_ptr, const int32_t* col, const float* s_src, const float* s_dst,
                                  float negative_slope, const float* h, int64_t ldh, int32_t feat, int64_t n_rows,
                                  const float* bias, const float* z, int64_t ldz, float beta, float* y, int64_t ldy,
                                  pgsd_stream_t stream) {
  PGSD_REQUIRE(n_rows >= 0 && feat >= 0, "gat_aggregate: negative size");
  if (n_rows == 0 || feat == 0) return PGSD_OK;
  PGSD_REQUIRE(row_ptr && s_src && s_dst && h && y, "gat_aggregate: null pointer");
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  PGSD_REQUIRE(feat % 4 == 0 && feat <= 128, "gat_aggregate: feat must be a multiple of 4, at most 128 (got %d)", feat);
  PGSD_REQUIRE(al16(h) && al16(y) && ldh % 4 == 0 && ldy % 4 == 0 && (bias == nullptr || al16(bias)) &&
                   (z == nullptr || (al16(z) && ldz % 4 == 0)),
               "gat_aggregate: rows must be 16-byte aligned");
  GatParams p{};
  p.n_rows = n_rows, p.feat = feat, p.lpr_active = feat / 4, p.slope = negative_slope, p.beta = beta;
  p.row_ptr = row_ptr, p.col = col, p.s_src = s_src, p.s_dst = s_dst, p.h = h, p.ldh = ldh, p.bias = bias;
  p.z = z, p.ldz = ldz, p.y = y, p.ldy = ldy;
  int lpr = 1;
  while (lpr < p.lpr_active) lpr <<= 1;
  int64_t grid = ceil_div<int64_t>(n_rows, 8 * (32 / lpr));
  if (grid > int64_t(sm_count()) * 5) grid = int64_t(sm_count()) * 5;   // one wave: 48 registers -> 5 CTAs per SM
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (lpr) {
    case 1: gat_aggregate_kernel<1><<<unsigned(grid), 256, 0, st>>>(p); break;
    case 2: gat_aggregate_kernel<2><<<unsigned(grid), 256, 0, st>>>(p); break;
    case 4: gat_aggregate_kernel<4><<<unsigned(grid), 256, 0, st>>>(p); break;
    case 8: gat_aggregate_kernel<8><<<unsigned(grid), 256, 0, st>>>(p); break;
    case 16: gat_aggregate_kernel<16><<<unsigned(grid), 256, 0, st>>>(p); break;
    default: gat_aggregate_kernel<32><<<unsigned(grid), 256, 0, st>>>(p); break;
  }
  PGSD_LAUNCH_CHECK("gat_aggregate_kernel");
  return PGSD_OK;
}


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
