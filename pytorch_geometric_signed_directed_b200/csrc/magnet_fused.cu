// One MagNetConv / MSConv layer (Chebyshev order K = 1) in ONE persistent kernel: the sparse
// aggregation of spmm.cu and the tcgen05 transform of dense_tc.cu fused through shared memory, so
// T = L~ x (2 x [N, 64] fp32 = 512 MB at the north-star size) is never written to or re-read from HBM.
//
//   T_k[r]   = diag_k[r] x_k[r] + sum_{e in row r} val_k[e] x_k[col[e]]          k = real, imag
//   A = x_real W0 + T_real W1,  B = x_imag W0 + T_imag W1
//   out_real = A - B + b,  out_imag = A + B + b                     (MagNetConv.py:185-249)
//
// Roles inside a CTA (one CTA per SM, 128-row tiles, tile t of CTA b = rows of tile b + t*grid):
//   * producer warps (PW of them): the group-per-row aggregation of spmm_groups_kernel -- a 16-lane
//     group owns a destination row, 128-bit gathers with U*2 loads in flight per lane, next index
//     batch prefetched behind the first gathers.  Rows are handed out by a shared-memory ticket
//     counter (dynamic balance inside the CTA); a finished row (T_real, T_imag: 2 x 256 B) is stored
//     into slot (tile & 1) of a two-slot ring and the group arrives on the slot's `full` mbarrier
//     (128 arrivals = one tile).  A group never spins: if the slot is still being read (it would
//     have to be two tiles ahead of the consumer) it idles with its sums in registers while the
//     other group of its warp keeps going.
//   * consumer warpgroup (warps 0-3): per tile, eight [128 x 32] fp32 operand chunks -- x_real,
//     x_imag straight from global (coalesced), T_real, T_imag from the ring -- are split into TF32
//     hi/lo, stored into a K-major SWIZZLE_128B staging buffer, and one thread issues the 3xTF32
//     tcgen05.mma triple (hi*hi + lo*hi + hi*lo, M=128, N=64, K=8) against the weight images that
//     stay resident in shared memory; accumulators A and B live in 128 TMEM columns; the epilogue
//     reads them with tcgen05.ld (warp q <-> TMEM lanes 32q..32q+31), mixes real/imag, adds the bias,
//     applies the optional complex-ReLU mask and streams the two output rows.
// The tile period (~60 us of gather traffic per SM) dwarfs the consumer's ~10 us, so the transform
// rides along for free and the kernel's HBM traffic is gathers + x + outputs only.
#include "tc_common.cuh"

namespace pgsd {
namespace fused {
using namespace tc;

constexpr int F_IN = 64;
constexpr int N_OUT = 64;
constexpr int TILE_M = 128;
constexpr int CW = 4;                                   // consumer warps = warps 0..3
constexpr int LPR = 16, NOPS = 2;
constexpr int ROW_BYTES = F_IN * 4;
constexpr int STAGE_HALF = TILE_M * 128;                // one [128 x 32] fp32 SW128 image
constexpr int STAGING_BYTES = 2 * STAGE_HALF;           // hi + lo
constexpr int W_HALF = N_OUT * 128;
constexpr int SLAB_BYTES = 2 * W_HALF;                  // hi + lo of one [32 x 64] weight slab
constexpr int N_SLABS = 4;                              // W0[k<32], W0[k>=32], W1[k<32], W1[k>=32]
constexpr int RING_OP_BYTES = TILE_M * ROW_BYTES;
constexpr int RING_SLOT_BYTES = NOPS * RING_OP_BYTES;
constexpr int RING_SLOTS = 2;
constexpr int SMEM_BYTES = 1024 + STAGING_BYTES + N_SLABS * SLAB_BYTES + RING_SLOTS * RING_SLOT_BYTES + 128;
constexpr uint32_t TMEM_COLS = 128;                     // A: columns [0, 64), B: [64, 128)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | (uint32_t(N_OUT >> 3) << 17) |
                           (uint32_t(TILE_M >> 4) << 24);   // F32 accum, TF32 x TF32, K-major, N=64, M=128

struct Params {
  int64_t n_rows, n_tiles;
  const int32_t* row_ptr;
  const int32_t* col;
  const float* val[2];
  const float* diag[2];
  float diag_const[2];
  const char* x[2];
  int64_t ldx_bytes[2];
  const float* w[2];
  int64_t ldw_k[2], ldw_n[2];
  const float* bias;
  char* y[2];
  int64_t ldy_bytes[2];
  int32_t relu_mode;
};

__device__ __forceinline__ void consumer_bar() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// wait that may legitimately last a whole tile period (and much longer under a sanitizer / profiler
// replay); still bounded so that a protocol bug traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait_long(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    __nanosleep(64);
    if (++spins > (1u << 27)) __trap();
  }
}

template <int PW, int U>
__global__ void __launch_bounds__((PW + CW) * 32, 1) magnet_layer_fused_kernel(const __grid_constant__ Params p) {
  constexpr int THREADS = (PW + CW) * 32;
  constexpr unsigned FULL = 0xffffffffu;

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* staging = smem;
  uint8_t* w_smem = staging + STAGING_BYTES;
  uint8_t* ring = w_smem + N_SLABS * SLAB_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ring + RING_SLOTS * RING_SLOT_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 16, bar_mma = bar_full + 32;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < RING_SLOTS; ++s) {
      mbar_init(bar_full + 8 * s, TILE_M);     // one arrival per finished row
      mbar_init(bar_empty + 8 * s, 1);
    }
    mbar_init(bar_mma, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights -> K-major SW128 hi/lo images (element (n, k) in 16-byte unit (k/4) ^ (n&7) of row n)
  for (int idx = tid; idx < N_SLABS * 32 * N_OUT; idx += THREADS) {
    const int s = idx / (32 * N_OUT);
    const int rem = idx - s * (32 * N_OUT);
    const int k = rem / N_OUT, n = rem - k * N_OUT;
    const int wi = s >> 1, kg = (s & 1) * 32 + k;
    const float v = __ldg(p.w[wi] + kg * p.ldw_k[wi] + n * p.ldw_n[wi]);
    const int off = n * 128 + (((k >> 2) ^ (n & 7)) << 4) + (k & 3) * 4;
    const float hi = to_tf32(v), lo = to_tf32(v - hi);
    *reinterpret_cast<float*>(w_smem + s * SLAB_BYTES + off) = hi;
    *reinterpret_cast<float*>(w_smem + s * SLAB_BYTES + W_HALF + off) = lo;
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < CW) {
    // ================================================================== consumer warpgroup
    const int r0 = tid >> 3, u16 = tid & 7;
    const uint32_t stg_addr = smem_u32(staging), w_addr = smem_u32(w_smem);
    const uint64_t pol_stream = policy_evict_first();
    const int64_t first = blockIdx.x;
    const int my_tiles = first < p.n_tiles ? int((p.n_tiles - first + gridDim.x - 1) / gridDim.x) : 0;
    uint32_t chunks_done = 0;
    for (int j = 0; j < my_tiles; ++j) {
      const int64_t tile = first + int64_t(j) * gridDim.x;
      const int slot = j & 1;
      const uint8_t* ring_slot = ring + slot * RING_SLOT_BYTES;
#pragma unroll 1
      for (int c = 0; c < 8; ++c) {
        const int k = (c >> 1) & 1, half = c & 1;
        const bool from_ring = c >= 4;
        float4 v[8];
        if (!from_ring) {
          const char* base = p.x[k] + half * 128 + u16 * 16;
#pragma unroll
          for (int pp = 0; pp < 8; ++pp) {
            const int64_t row = tile * TILE_M + pp * 16 + r0;
            v[pp] = row < p.n_rows ? __ldg(reinterpret_cast<const float4*>(base + row * p.ldx_bytes[k]))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        } else {
          if (c == 4) mbar_wait_long(bar_full + 8 * slot, (j >> 1) & 1);
          const uint8_t* base = ring_slot + k * RING_OP_BYTES + half * 128 + u16 * 16;
#pragma unroll
          for (int pp = 0; pp < 8; ++pp)
            v[pp] = *reinterpret_cast<const float4*>(base + (pp * 16 + r0) * ROW_BYTES);
        }
        if (chunks_done > 0) mbar_wait(bar_mma, (chunks_done - 1) & 1);    // staging free again
#pragma unroll
        for (int pp = 0; pp < 8; ++pp) {
          const int rr = pp * 16 + r0;
          const uint32_t off = uint32_t(rr) * 128 + uint32_t((u16 ^ (rr & 7)) << 4);
          float4 vh, vl;
          vh.x = to_tf32(v[pp].x), vh.y = to_tf32(v[pp].y), vh.z = to_tf32(v[pp].z), vh.w = to_tf32(v[pp].w);
          vl.x = to_tf32(v[pp].x - vh.x), vl.y = to_tf32(v[pp].y - vh.y);
          vl.z = to_tf32(v[pp].z - vh.z), vl.w = to_tf32(v[pp].w - vh.w);
          *reinterpret_cast<float4*>(staging + off) = vh;
          *reinterpret_cast<float4*>(staging + STAGE_HALF + off) = vl;
        }
        fence_async_smem();
        consumer_bar();
        if (tid == 0) {
          if (c == 7) mbar_arrive(bar_empty + 8 * slot);       // every consumer thread has read the slot
          tc_fence_after();
          const uint32_t d = tmem_base + uint32_t(k) * N_OUT;
          const uint32_t a_hi = stg_addr, a_lo = stg_addr + STAGE_HALF;
          const uint32_t w_hi = w_addr + uint32_t((from_ring ? 2 : 0) + half) * SLAB_BYTES, w_lo = w_hi + W_HALF;
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint32_t ko = jj * 32;
            const uint32_t acc0 = (c < 4 && half == 0 && jj == 0) ? 0u : 1u;   // first chunk of A / B overwrites
            mma_tf32(d, make_desc(a_hi + ko), make_desc(w_hi + ko), IDESC, acc0);
            mma_tf32(d, make_desc(a_lo + ko), make_desc(w_hi + ko), IDESC, 1u);
            mma_tf32(d, make_desc(a_hi + ko), make_desc(w_lo + ko), IDESC, 1u);
          }
          tc_commit(bar_mma);
        }
        ++chunks_done;
      }
      // ---- epilogue: TMEM -> registers -> global
      mbar_wait(bar_mma, (chunks_done - 1) & 1);
      tc_fence_after();
      {
        const int64_t row = tile * TILE_M + warp * 32 + lane;
        const uint32_t taddr = tmem_base + (uint32_t(warp * 32) << 16);
#pragma unroll 1
        for (int cb = 0; cb < N_OUT / 16; ++cb) {
          float a[16], b[16];
          tmem_ld<16>(taddr + cb * 16, a);
          tmem_ld<16>(taddr + N_OUT + cb * 16, b);
          tmem_ld_wait();
          if (row < p.n_rows) {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const float bs = p.bias ? __ldg(p.bias + cb * 16 + i) : 0.f;
              const float o0 = (a[i] - b[i]) + bs, o1 = (a[i] + b[i]) + bs;
              const float m = (p.relu_mode == 1 && !(o0 >= 0.f)) ? 0.f : 1.f;
              a[i] = p.relu_mode == 1 ? o0 * m : o0;
              b[i] = p.relu_mode == 1 ? o1 * m : o1;
            }
            char* y0 = p.y[0] + row * p.ldy_bytes[0] + cb * 64;
            char* y1 = p.y[1] + row * p.ldy_bytes[1] + cb * 64;
#pragma unroll
            for (int i = 0; i < 16; i += 4) {
              st_stream_v4(y0 + i * 4, make_float4(a[i], a[i + 1], a[i + 2], a[i + 3]), pol_stream);
              st_stream_v4(y1 + i * 4, make_float4(b[i], b[i + 1], b[i + 2], b[i + 3]), pol_stream);
            }
          }
        }
      }
      tc_fence_before();
      consumer_bar();          // accumulators are free for the next tile's overwriting MMA
    }
  } else {
    // ====================================================================== producer warps
    const int g = lane >> 4, l = lane & 15;
    const unsigned gmask = 0xffffu << (16 * g);
    const char* xb[NOPS];
    uint32_t ldx32[NOPS];
#pragma unroll
    for (int k = 0; k < NOPS; ++k) xb[k] = p.x[k] + l * 16, ldx32[k] = uint32_t(p.ldx_bytes[k]);

    // ticket t -> row (t & 127) of this CTA's (t >> 7)-th tile; row_of() < 0: no tile left
    auto row_of = [&](int t) -> int {
      const int64_t tile = int64_t(blockIdx.x) + int64_t(t >> 7) * gridDim.x;
      return tile < p.n_tiles ? int(tile * TILE_M + (t & (TILE_M - 1))) : -1;
    };
    auto load_ptrs = [&](int r, int& s, int& e) {
      s = e = 0;
      if (r >= 0 && r < p.n_rows) s = __ldg(p.row_ptr + r), e = __ldg(p.row_ptr + r + 1);
    };
    auto load_batch = [&](int base, int end, int& c, float (&v)[NOPS]) {
      c = 0;
#pragma unroll
      for (int k = 0; k < NOPS; ++k) v[k] = 0.f;      // zero values keep lanes past the row end inert
      const int e = base + l;
      if (e < end) {
        c = ld_once_i32(p.col + e);
#pragma unroll
        for (int k = 0; k < NOPS; ++k) v[k] = ld_once_f32(p.val[k] + e);
      }
    };

    // Rows are dealt round-robin: group gi takes tickets gi, gi + NG, gi + 2 NG, ... of the CTA's row
    // sequence (like the grid-stride assignment of spmm_groups_kernel: equal row counts, row-length
    // variance averages out over the ~170 rows a group sees at the north-star size).
    constexpr int NG = PW * 2;
    int ct = (warp - CW) * 2 + g;      // current ticket; the next one is ct + NG
    int base, end, nstart, nend, c;
    float v[NOPS];
    bool live = row_of(ct) >= 0;       // the current ticket maps to a tile of this CTA
    load_ptrs(row_of(ct), base, end);
    load_batch(base, end, c, v);
    load_ptrs(row_of(ct + NG), nstart, nend);

    float acc[NOPS][4];
#pragma unroll
    for (int k = 0; k < NOPS; ++k)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
    int pending = 0;           // row finished, waiting for its ring slot; (c, v) already hold the next row's batch

    // Booleans are kept out of the gather section on purpose: with more than ~4 live predicates next
    // to the U "entry exists" predicates, ptxas recycles a predicate register between the gathers of a
    // batch, which forces the zero/select of an earlier gather (a wait on its data) in front of the
    // later loads (ncu source page, session 13).  The section below needs only ok[u] and the loop test.
    while (__any_sync(FULL, live)) {
      const int cnt = (live && !pending) ? min(LPR, end - base) : 0;        // <= 0: nothing to gather
      // index range to prefetch in this iteration: the row's next batch, else the next row's first one
      int pf_b = 0, pf_e = 0;
      if (live && !pending) {
        if (base + LPR < end) pf_b = base + LPR, pf_e = end;
        else pf_b = nstart, pf_e = nend;
      }
      int nc = 0;
      float nv[NOPS] = {0.f, 0.f};
      const bool ran = __any_sync(FULL, cnt > 0);
      // Every gather of a pass is issued UNCONDITIONALLY and its FFMAs are predicated instead: a slot
      // past the end of the batch re-reads the batch's last neighbour row (an L2 hit, a different row
      // for every group -- not one hot sector), an idle group reads a row of its next batch.  With
      // predicated loads ptxas zero-fills through MOV/selects that wait on the data in the middle
      // of the gather sequence, which delayed the last gather and the index prefetch of every pass
      // by one memory round trip (ncu source page, sessions 13-14).
      const int last = cnt > 0 ? cnt - 1 : 0;
      for (int jj = 0; __any_sync(FULL, jj < cnt); jj += U) {
        float d[NOPS][U][4];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const int cc = __shfl_sync(FULL, c, min(jj + u, last), LPR);
#pragma unroll
          for (int k = 0; k < NOPS; ++k) ld_gather_v4_plain_to(row_addr(xb[k], cc, ldx32[k]), d[k][u]);
        }
        // the next index batch goes out right behind the gathers
        if (jj == 0) load_batch(pf_b, pf_e, nc, nv);
#pragma unroll
        for (int u = 0; u < U; ++u)
#pragma unroll
          for (int k = 0; k < NOPS; ++k) {
            const float t = __shfl_sync(FULL, v[k], (jj + u) & (LPR - 1), LPR);
            if (jj + u < cnt) {
#pragma unroll
              for (int i = 0; i < 4; ++i) acc[k][i] = fmaf(t, d[k][u][i], acc[k][i]);
            }
          }
      }
      if (!ran) load_batch(pf_b, pf_e, nc, nv);       // no group of this warp had entries
      if (live && !pending) {
        c = nc;
#pragma unroll
        for (int k = 0; k < NOPS; ++k) v[k] = nv[k];
        if (base + LPR < end) base += LPR;
        else pending = 1;
      }
      if (pending) {
        const int cj = ct >> 7, slot = cj & 1;
        int free_ = 1;
        if (cj >= RING_SLOTS) {     // the consumer must have drained tile cj - 2 out of this slot
          if (l == 0) free_ = mbar_try(bar_empty + 8 * slot, uint32_t((cj >> 1) - 1) & 1u) ? 1 : 0;
          free_ = __shfl_sync(gmask, free_, 0, LPR);
        }
        if (free_) {
          const int crow = row_of(ct);
          const bool row_ok = crow < p.n_rows;
          uint8_t* dst = ring + slot * RING_SLOT_BYTES + (ct & (TILE_M - 1)) * ROW_BYTES + l * 16;
#pragma unroll
          for (int k = 0; k < NOPS; ++k) {
            const bool has_diag = p.diag[k] != nullptr;
            const float dg = !row_ok ? 0.f : (has_diag ? __ldg(p.diag[k] + crow) : p.diag_const[k]);
            if (row_ok && (has_diag || dg != 0.f)) {
              const float4 xr = __ldg(reinterpret_cast<const float4*>(row_addr(xb[k], crow, ldx32[k])));
              acc[k][0] = fmaf(dg, xr.x, acc[k][0]);
              acc[k][1] = fmaf(dg, xr.y, acc[k][1]);
              acc[k][2] = fmaf(dg, xr.z, acc[k][2]);
              acc[k][3] = fmaf(dg, xr.w, acc[k][3]);
            }
            *reinterpret_cast<float4*>(dst + k * RING_OP_BYTES) = make_float4(acc[k][0], acc[k][1], acc[k][2], acc[k][3]);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[k][i] = 0.f;
          }
          __syncwarp(gmask);
          if (l == 0) mbar_arrive(bar_full + 8 * slot);
          // advance to the row whose pointers and first batch are already here
          ct += NG, base = nstart, end = nend;
          live = row_of(ct) >= 0;
          pending = 0;
          load_ptrs(row_of(ct + NG), nstart, nend);
        }
      }
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_slot), "r"(TMEM_COLS)
                 : "memory");
  }
}

template <int PW, int U>
static int launch(const Params& p, cudaStream_t st) {
  auto kern = magnet_layer_fused_kernel<PW, U>;
  PGSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  int64_t grid = sm_count();
  if (grid > p.n_tiles) grid = p.n_tiles;
  kern<<<dim3(unsigned(grid)), dim3((PW + CW) * 32), SMEM_BYTES, st>>>(p);
  PGSD_LAUNCH_CHECK("magnet_layer_fused_kernel");
  return PGSD_OK;
}

static inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace fused
}  // namespace pgsd

using namespace pgsd;

extern "C" int pgsd_magnet_fused_supported(int32_t feat_in, int32_t feat_out, int32_t dtype) {
  return (feat_in == fused::F_IN && feat_out == fused::N_OUT && dtype == PGSD_F32) ? 1 : 0;
}

extern "C" size_t pgsd_sizeof_magnet_fused_args(void) { return sizeof(pgsd_magnet_fused_args); }

extern "C" int pgsd_magnet_layer_fused(const pgsd_magnet_fused_args* a, pgsd_stream_t stream) {
  using namespace fused;
  PGSD_REQUIRE(a != nullptr, "magnet_fused: args is null");
  PGSD_REQUIRE(pgsd_magnet_fused_supported(a->feat_in, a->feat_out, PGSD_F32),
               "magnet_fused: only feat_in = %d, feat_out = %d fp32 is supported (got %d -> %d)", F_IN, N_OUT,
               a->feat_in, a->feat_out);
  PGSD_REQUIRE(a->n_rows >= 0 && a->n_rows < (int64_t(1) << 31) - 2 * TILE_M, "magnet_fused: n_rows out of the int32 plan range");
  if (a->n_rows == 0) return PGSD_OK;
  PGSD_REQUIRE(a->row_ptr != nullptr, "magnet_fused: row_ptr is null");
  Params p{};
  p.n_rows = a->n_rows;
  p.n_tiles = (a->n_rows + TILE_M - 1) / TILE_M;
  p.row_ptr = a->row_ptr;
  p.col = a->col;
  p.bias = a->bias;
  p.relu_mode = a->relu_mode;
  for (int k = 0; k < 2; ++k) {
    // val / col may be null for a plan without entries (their loads are predicated on row_ptr)
    PGSD_REQUIRE(a->x[k] && a->y[k] && a->w[k], "magnet_fused: x/y/w[%d] is null", k);
    PGSD_REQUIRE(a->ldx[k] >= F_IN && a->ldy[k] >= N_OUT, "magnet_fused: leading dim < feature width");
    PGSD_REQUIRE(a->ldx[k] * 4 < (int64_t(1) << 32), "magnet_fused: row stride of x must be below 4 GiB");
    PGSD_REQUIRE(al16(a->x[k]) && al16(a->y[k]) && (a->ldx[k] * 4) % 16 == 0 && (a->ldy[k] * 4) % 16 == 0,
                 "magnet_fused: x / y rows must be 16-byte aligned");
    p.val[k] = a->val[k];
    p.diag[k] = a->diag[k];
    p.diag_const[k] = a->diag_const[k];
    p.x[k] = reinterpret_cast<const char*>(a->x[k]);
    p.ldx_bytes[k] = a->ldx[k] * 4;
    p.w[k] = a->w[k];
    p.ldw_k[k] = a->ldw_k[k];
    p.ldw_n[k] = a->ldw_n[k];
    p.y[k] = reinterpret_cast<char*>(a->y[k]);
    p.ldy_bytes[k] = a->ldy[k] * 4;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  switch (a->variant) {
    case 1: return launch<20, 4>(p, st);
    case 2: return launch<20, 2>(p, st);
    case 3: return launch<24, 2>(p, st);
    default: return launch<16, 4>(p, st);   // 640 threads -> 96 registers: the 8 gathers of a batch get 8 register quads
  }
}
