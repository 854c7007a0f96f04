// tcgen05 (5th-gen tensor core) path of the dense feature transform, fp32 in / fp32 out.
//
//   acc_g[128 x N] (TMEM, fp32) = sum over 32-wide K chunks  X_chunk[128 x 32] @ W_slab[32 x N]
//   out_real = acc_0 - acc_1 + b, out_imag = acc_0 + acc_1 + b       (combine, MagNetConv.py:242-247)
//
// fp32 parity through tensor cores: 3xTF32 error-compensated product.  Every operand is split
// in registers into hi = rna_tf32(v) and lo = rna_tf32(v - hi) and three kind::tf32 MMAs
// (hi*hi + lo*hi + hi*lo) accumulate into the same TMEM tile; the dropped lo*lo term is
// ~2^-22 relative, i.e. the result is fp32-class like the reference's sgemm (allow_tf32=False).
//
// Structure (one persistent CTA per SM, 512 threads, every thread plays every role):
//   * prologue: tcgen05.alloc of GROUPS*N TMEM columns; all weight slabs split into hi/lo and
//     written once to shared memory in the K-major SWIZZLE_128B canonical layout;
//   * per 128-row tile, per K chunk: coalesced 128-bit global loads issued PF chunks ahead
//     (register ring, ~64 KB in flight per SM) -> hi/lo split -> st.shared into a 2-stage
//     SW128 A buffer -> fence.proxy.async + bar -> one thread issues 12 tcgen05.mma
//     (M=128, N, K=8) and a tcgen05.commit onto the stage's mbarrier;
//   * epilogue: tcgen05.ld 32x32b (warp w reads TMEM lane quarter w%4, column block w/4),
//     real/imag mix + bias (+ complex ReLU mask), 128-bit streaming stores.
// The tensor work is ~1.6 us per tile against ~3 us of HBM time: the kernel is a streaming
// kernel whose math rides on the tensor pipe for free ("never the bound").
#include "tc_common.cuh"

namespace pgsd {
namespace tc {

constexpr int THREADS = 512;
constexpr int TILE_M = 128;
constexpr int STAGES = 2;
constexpr int STAGE_BYTES = 2 * TILE_M * 128;   // hi + lo
constexpr int PF = 4;                    // chunks in flight per thread (2 x LDG.128 each)
constexpr int MAX_CHUNKS = 2 * PGSD_DENSE_MAX_TERMS;
constexpr int MAX_SLABS = 16;

struct Params {
  int64_t n_rows;
  int32_t n_chunks, n_slabs, relu_mode, n_total;   // n_total: full output width (N tiles over grid.y)
  int32_t stages, pad;                             // smem pipeline depth of the warp-specialised kernel
  const char* x[MAX_CHUNKS];      // term base + k0 (bytes already applied)
  int64_t ldx_bytes[MAX_CHUNKS];
  int8_t group[MAX_CHUNKS];
  int8_t slab[MAX_CHUNKS];
  int8_t first[MAX_CHUNKS];       // first chunk of its group inside a tile -> overwrite TMEM
  int8_t kvalid[MAX_CHUNKS];      // valid k in this chunk (fp32: <= 32, bf16: <= 64)
  const float* w[MAX_SLABS];      // weight base + k0 * ldw_k
  int64_t ldw_k[MAX_SLABS], ldw_n[MAX_SLABS];
  int8_t wk[MAX_SLABS];           // valid k rows of the slab
  const float* bias;
  char* y[2];
  int64_t ldy_bytes[2];
};

// ------------------------------------------------------------------------------------ kernel
// BF16 = false: fp32 operands, 3xTF32 (hi/lo split), 32 k per 128-byte chunk row.
// BF16 = true : bf16 operands fed to kind::f16 directly, 64 k per 128-byte chunk row, bf16 output.
// grid.y tiles the output width in N_OUT-column tiles (weights / bias / outputs offset by n0).
template <int N_OUT, int GROUPS, bool BF16, int PFK>
__global__ void __launch_bounds__(THREADS, 1) dense_tc_kernel(const __grid_constant__ Params p) {
  constexpr int CPW = N_OUT / 4;                 // output columns per epilogue warp
  constexpr int ES = BF16 ? 2 : 4;               // operand element size
  constexpr int EPC = 16 / ES;                   // elements per 16-byte unit
  constexpr int CK = 128 / ES;                   // k per chunk
  constexpr int HALF = N_OUT * 128;              // one [N_OUT x 128 B] weight image
  constexpr int SLAB_BYTES = BF16 ? HALF : 2 * HALF;
  constexpr int UMMA_K_BYTES = 32;               // 8 tf32 or 16 bf16
  constexpr uint32_t TMEM_COLS = (GROUPS * N_OUT <= 32) ? 32 : (GROUPS * N_OUT <= 64) ? 64
                               : (GROUPS * N_OUT <= 128) ? 128 : 256;
  // instruction descriptor: c=F32 [4,6), a/b format [7,10)/[10,13) (TF32 = 2, BF16 = 1), K-major
  // both, N>>3 [17,23), M>>4 [24,29)
  constexpr uint32_t FMT = BF16 ? 1u : 2u;
  constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | (uint32_t(N_OUT >> 3) << 17) |
                             (uint32_t(TILE_M >> 4) << 24);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem is only guaranteed 16-byte aligned: round up to the 1024 B the swizzle needs
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* a_stage = smem;
  uint8_t* w_smem = smem + STAGES * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + p.n_slabs * SLAB_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + STAGES + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * N_OUT;             // first output column of this CTA
  const uint32_t a_addr = smem_u32(a_stage), w_addr = smem_u32(w_smem);
  const uint32_t bar_addr = smem_u32(bars);
  const uint64_t pol_stream = policy_evict_first();

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s <= STAGES; ++s) mbar_init(bar_addr + 8 * s, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights -> K-major SW128 images: element (n, k) lives in 16-byte unit (k / EPC) ^ (n & 7) of
  // row n (128 B per row).  fp32: hi image then lo image; bf16: one rounded image.
  for (int idx = tid; idx < p.n_slabs * CK * N_OUT; idx += THREADS) {
    const int s = idx / (CK * N_OUT);
    const int rem = idx - s * (CK * N_OUT);
    const int k = rem / N_OUT, n = rem - k * N_OUT;
    float v = 0.f;
    if (k < p.wk[s] && n0 + n < p.n_total) v = __ldg(p.w[s] + k * p.ldw_k[s] + (n0 + n) * p.ldw_n[s]);
    const int off = n * 128 + (((k / EPC) ^ (n & 7)) << 4) + (k % EPC) * ES;
    if constexpr (BF16) {
      *reinterpret_cast<__nv_bfloat16*>(w_smem + s * SLAB_BYTES + off) = __float2bfloat16_rn(v);
    } else {
      const float hi = to_tf32(v), lo = to_tf32(v - hi);
      *reinterpret_cast<float*>(w_smem + s * SLAB_BYTES + off) = hi;
      *reinterpret_cast<float*>(w_smem + s * SLAB_BYTES + HALF + off) = lo;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // loader role: rows r0 = tid/8 and r0 + 64 of the tile, 16-byte unit c16 = tid%8 of the row
  const int r0 = tid >> 3, c16 = tid & 7;
  const uint32_t st_off0 = uint32_t(r0) * 128 + uint32_t((c16 ^ (r0 & 7)) << 4);
  const uint32_t st_off1 = st_off0 + 64 * 128;   // (r0 + 64) & 7 == r0 & 7

  const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;
  float4 buf[PFK][2];
  auto issue = [&](int slot, int64_t tile, int c) {
    const bool kin = (c16 * EPC) < p.kvalid[c];
    const char* base = p.x[c] + c16 * 16;
    const int64_t ra = tile * TILE_M + r0, rb = ra + 64;
    buf[slot][0] = (tile < n_tiles && kin && ra < p.n_rows)
                       ? __ldg(reinterpret_cast<const float4*>(base + ra * p.ldx_bytes[c]))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
    buf[slot][1] = (tile < n_tiles && kin && rb < p.n_rows)
                       ? __ldg(reinterpret_cast<const float4*>(base + rb * p.ldx_bytes[c]))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
  };
#pragma unroll
  for (int u = 0; u < PFK; ++u) {
    buf[u][0] = buf[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (u < p.n_chunks) issue(u, blockIdx.x, u);
  }

  uint32_t uses = 0;   // chunks staged so far (CTA-uniform)
  uint32_t tiles_done = 0;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int cb = 0; cb < p.n_chunks; cb += PFK) {
#pragma unroll
      for (int u = 0; u < PFK; ++u) {
        const int c = cb + u;
        if (c < p.n_chunks) {
          const uint32_t stage = uses % STAGES;
          if (uses >= STAGES) mbar_wait(bar_addr + 8 * stage, ((uses / STAGES) - 1) & 1);
          uint8_t* hi = a_stage + stage * STAGE_BYTES;
          uint8_t* lo = hi + TILE_M * 128;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const float4 v = buf[u][h];
            const uint32_t off = h ? st_off1 : st_off0;
            if constexpr (BF16) {
              *reinterpret_cast<float4*>(hi + off) = v;      // 8 bf16, already in operand format
            } else {                                          // split the fp32 values into tf32 hi/lo
              float4 vh, vl;
              vh.x = to_tf32(v.x), vh.y = to_tf32(v.y), vh.z = to_tf32(v.z), vh.w = to_tf32(v.w);
              vl.x = to_tf32(v.x - vh.x), vl.y = to_tf32(v.y - vh.y);
              vl.z = to_tf32(v.z - vh.z), vl.w = to_tf32(v.w - vh.w);
              *reinterpret_cast<float4*>(hi + off) = vh;
              *reinterpret_cast<float4*>(lo + off) = vl;
            }
          }
          // refill this register slot with the next chunk that maps to it
          {
            int nc = c + PFK;
            int64_t nt = tile;
            if (nc >= p.n_chunks) nc = u, nt = tile + gridDim.x;
            issue(u, nt, nc);
          }
          fence_async_smem();
          __syncthreads();
          if (tid == 0) {
            tc_fence_after();
            const uint32_t d = tmem_base + uint32_t(p.group[c]) * N_OUT;
            const uint32_t a_hi = a_addr + stage * STAGE_BYTES, a_lo = a_hi + TILE_M * 128;
            const uint32_t w_hi = w_addr + uint32_t(p.slab[c]) * SLAB_BYTES, w_lo = w_hi + HALF;
#pragma unroll
            for (int j = 0; j < 128 / UMMA_K_BYTES; ++j) {
              const uint32_t ko = j * UMMA_K_BYTES;   // one MMA's K extent inside the 128-byte swizzle row
              const uint32_t acc0 = (p.first[c] && j == 0) ? 0u : 1u;
              if constexpr (BF16) {
                mma_bf16(d, make_desc(a_hi + ko), make_desc(w_hi + ko), IDESC, acc0);
              } else {
                mma_tf32(d, make_desc(a_hi + ko), make_desc(w_hi + ko), IDESC, acc0);
                mma_tf32(d, make_desc(a_lo + ko), make_desc(w_hi + ko), IDESC, 1u);
                mma_tf32(d, make_desc(a_hi + ko), make_desc(w_lo + ko), IDESC, 1u);
              }
            }
            tc_commit(bar_addr + 8 * stage);                      // stage may be overwritten
            if (c == p.n_chunks - 1) tc_commit(bar_addr + 8 * STAGES);   // accumulators complete
          }
          ++uses;
        }
      }
    }

    // ---- epilogue: TMEM -> registers -> global
    mbar_wait(bar_addr + 8 * STAGES, tiles_done & 1);
    tc_fence_after();
    {
      const int q = warp & 3, cblk = warp >> 2;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + uint32_t(cblk * CPW);
      float a[CPW], b[CPW];
      tmem_ld<CPW>(taddr, a);
      if (GROUPS == 2) tmem_ld<CPW>(taddr + N_OUT, b);
      tmem_ld_wait();
      const int64_t row = tile * TILE_M + q * 32 + lane;
      const int ncol = n0 + cblk * CPW;            // first global output column of this thread
      if (row < p.n_rows && ncol < p.n_total) {
        float o0[CPW], o1[CPW];
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
          const float bs = p.bias ? __ldg(p.bias + ncol + i) : 0.f;
          if (GROUPS == 2) {
            o0[i] = (a[i] - b[i]) + bs;
            o1[i] = (a[i] + b[i]) + bs;
            if (p.relu_mode == 1) {
              const float m = o0[i] >= 0.f ? 1.f : 0.f;
              o0[i] *= m, o1[i] *= m;
            }
          } else {
            o0[i] = a[i] + bs;
            if (p.relu_mode == 2) o0[i] = tanhf(o0[i]);
          }
        }
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
          const float* o = g ? o1 : o0;
          char* yp = p.y[g] + row * p.ldy_bytes[g] + int64_t(ncol) * ES;
          if constexpr (BF16) {
#pragma unroll
            for (int i = 0; i < CPW; i += 8) {
              float4 pk;
              pk.x = __uint_as_float(pack_bf16(o[i], o[i + 1]));
              pk.y = __uint_as_float(pack_bf16(o[i + 2], o[i + 3]));
              pk.z = __uint_as_float(pack_bf16(o[i + 4], o[i + 5]));
              pk.w = __uint_as_float(pack_bf16(o[i + 6], o[i + 7]));
              st_stream_v4(yp + i * 2, pk, pol_stream);
            }
          } else {
#pragma unroll
            for (int i = 0; i < CPW; i += 4)
              st_stream_v4(yp + i * 4, make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]), pol_stream);
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();     // accumulators are free for the next tile's first MMA
    ++tiles_done;
  }

  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------ warp-specialised
// Same math, decoupled roles (the synchronous kernel above spends ~30 % of its time at the per-chunk
// __syncthreads waiting for the one thread that builds descriptors and issues the MMAs):
//   warps 0..15 : loaders (global -> registers -> hi/lo split -> swizzled st.shared) and epilogue
//   warp  16    : MMA issuer (one elected lane): waits full[stage], issues, commits empty[stage]
// Hand-offs are mbarriers only; TMEM accumulators are double-buffered so tile t+1's MMAs overlap
// tile t's epilogue.
constexpr int WS_LOADER_WARPS = 16;
constexpr int WS_THREADS = (WS_LOADER_WARPS + 1) * 32;
constexpr int WS_MAX_STAGES = 4;


template <int N_OUT, int GROUPS, bool BF16>
__global__ void __launch_bounds__(WS_THREADS, 1) dense_tc_ws_kernel(const __grid_constant__ Params p) {
  constexpr int CPW = N_OUT / 4;
  constexpr int ES = BF16 ? 2 : 4;
  constexpr int EPC = 16 / ES;
  constexpr int CK = 128 / ES;
  constexpr int HALF = N_OUT * 128;
  constexpr int SLAB_BYTES = BF16 ? HALF : 2 * HALF;
  constexpr int UMMA_K_BYTES = 32;
  constexpr int ACC_COLS = GROUPS * N_OUT;
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64
                               : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  constexpr uint32_t FMT = BF16 ? 1u : 2u;
  constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | (uint32_t(N_OUT >> 3) << 17) |
                             (uint32_t(TILE_M >> 4) << 24);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages;                                // pipeline stages chosen by the host (2..4)
  uint8_t* a_stage = smem;
  uint8_t* w_smem = smem + S * STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + p.n_slabs * SLAB_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * WS_MAX_STAGES + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * N_OUT;
  const uint32_t a_addr = smem_u32(a_stage), w_addr = smem_u32(w_smem);
  const uint32_t bar_full = smem_u32(bars), bar_empty = bar_full + 8 * WS_MAX_STAGES;
  const uint32_t bar_acc_full = bar_empty + 8 * WS_MAX_STAGES, bar_acc_empty = bar_acc_full + 16;
  const uint64_t pol_stream = policy_evict_first();

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < WS_MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, WS_LOADER_WARPS);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, WS_LOADER_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int idx = tid; idx < p.n_slabs * CK * N_OUT; idx += WS_THREADS) {
    const int s = idx / (CK * N_OUT);
    const int rem = idx - s * (CK * N_OUT);
    const int k = rem / N_OUT, n = rem - k * N_OUT;
    float v = 0.f;
    if (k < p.wk[s] && n0 + n < p.n_total) v = __ldg(p.w[s] + k * p.ldw_k[s] + (n0 + n) * p.ldw_n[s]);
    const int off = n * 128 + (((k / EPC) ^ (n & 7)) << 4) + (k % EPC) * ES;
    if constexpr (BF16) {
      *reinterpret_cast<__nv_bfloat16*>(w_smem + s * SLAB_BYTES + off) = __float2bfloat16_rn(v);
    } else {
      const float hi = to_tf32(v), lo = to_tf32(v - hi);
      *reinterpret_cast<float*>(w_smem + s * SLAB_BYTES + off) = hi;
      *reinterpret_cast<float*>(w_smem + s * SLAB_BYTES + HALF + off) = lo;
    }
  }
  fence_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;

  if (warp == WS_LOADER_WARPS) {
    // ===================================================================== MMA issuer
    uint32_t uses = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      if (it >= 2) mbar_wait(bar_acc_empty + 8 * acc, ((it >> 1) - 1) & 1);   // epilogue drained this buffer
      for (int c = 0; c < p.n_chunks; ++c, ++uses) {
        const uint32_t stage = uses % S;
        mbar_wait(bar_full + 8 * stage, (uses / S) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t d = tmem_base + acc * ACC_COLS + uint32_t(p.group[c]) * N_OUT;
          const uint32_t a_hi = a_addr + stage * STAGE_BYTES, a_lo = a_hi + TILE_M * 128;
          const uint32_t w_hi = w_addr + uint32_t(p.slab[c]) * SLAB_BYTES, w_lo = w_hi + HALF;
#pragma unroll
          for (int j = 0; j < 128 / UMMA_K_BYTES; ++j) {
            const uint32_t ko = j * UMMA_K_BYTES;
            const uint32_t acc0 = (p.first[c] && j == 0) ? 0u : 1u;
            if constexpr (BF16) {
              mma_bf16(d, make_desc(a_hi + ko), make_desc(w_hi + ko), IDESC, acc0);
            } else {
              mma_tf32(d, make_desc(a_hi + ko), make_desc(w_hi + ko), IDESC, acc0);
              mma_tf32(d, make_desc(a_lo + ko), make_desc(w_hi + ko), IDESC, 1u);
              mma_tf32(d, make_desc(a_hi + ko), make_desc(w_lo + ko), IDESC, 1u);
            }
          }
          tc_commit(bar_empty + 8 * stage);
          if (c == p.n_chunks - 1) tc_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
      }
    }
  } else {
    // ============================================================ loaders + epilogue warps
    const int r0 = tid >> 3, c16 = tid & 7;
    const uint32_t st_off0 = uint32_t(r0) * 128 + uint32_t((c16 ^ (r0 & 7)) << 4);
    const uint32_t st_off1 = st_off0 + 64 * 128;
    float4 buf[PF][2];
    auto issue = [&](int slot, int64_t tile, int c) {
      const bool kin = (c16 * EPC) < p.kvalid[c];
      const char* base = p.x[c] + c16 * 16;
      const int64_t ra = tile * TILE_M + r0, rb = ra + 64;
      buf[slot][0] = (tile < n_tiles && kin && ra < p.n_rows)
                         ? __ldg(reinterpret_cast<const float4*>(base + ra * p.ldx_bytes[c]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      buf[slot][1] = (tile < n_tiles && kin && rb < p.n_rows)
                         ? __ldg(reinterpret_cast<const float4*>(base + rb * p.ldx_bytes[c]))
                         : make_float4(0.f, 0.f, 0.f, 0.f);
    };
#pragma unroll
    for (int u = 0; u < PF; ++u) {
      buf[u][0] = buf[u][1] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (u < p.n_chunks) issue(u, blockIdx.x, u);
    }
    uint32_t uses = 0, it = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      for (int cb = 0; cb < p.n_chunks; cb += PF) {
#pragma unroll
        for (int u = 0; u < PF; ++u) {
          const int c = cb + u;
          if (c < p.n_chunks) {
            const uint32_t stage = uses % S;
            if (uses >= uint32_t(S)) mbar_wait(bar_empty + 8 * stage, ((uses / S) - 1) & 1);
            uint8_t* hi = a_stage + stage * STAGE_BYTES;
            uint8_t* lo = hi + TILE_M * 128;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const float4 v = buf[u][h];
              const uint32_t off = h ? st_off1 : st_off0;
              if constexpr (BF16) {
                *reinterpret_cast<float4*>(hi + off) = v;
              } else {
                float4 vh, vl;
                vh.x = to_tf32(v.x), vh.y = to_tf32(v.y), vh.z = to_tf32(v.z), vh.w = to_tf32(v.w);
                vl.x = to_tf32(v.x - vh.x), vl.y = to_tf32(v.y - vh.y);
                vl.z = to_tf32(v.z - vh.z), vl.w = to_tf32(v.w - vh.w);
                *reinterpret_cast<float4*>(hi + off) = vh;
                *reinterpret_cast<float4*>(lo + off) = vl;
              }
            }
            {
              int nc = c + PF;
              int64_t nt = tile;
              if (nc >= p.n_chunks) nc = u, nt = tile + gridDim.x;
              issue(u, nt, nc);
            }
            fence_async_smem();                 // generic-proxy stores -> visible to the tensor core
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_full + 8 * stage);
            ++uses;
          }
        }
      }
      // ---- epilogue of this tile (the next tile's loads are already in flight in `buf`)
      const uint32_t acc = it & 1;
      mbar_wait(bar_acc_full + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const int q = warp & 3, cblk = warp >> 2;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * ACC_COLS + uint32_t(cblk * CPW);
      float a[CPW], b[CPW];
      tmem_ld<CPW>(taddr, a);
      if (GROUPS == 2) tmem_ld<CPW>(taddr + N_OUT, b);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);     // TMEM buffer free for tile it+2
      const int64_t row = tile * TILE_M + q * 32 + lane;
      const int ncol = n0 + cblk * CPW;
      if (row < p.n_rows && ncol < p.n_total) {
#pragma unroll
        for (int i = 0; i < CPW; ++i) {
          const float bs = p.bias ? __ldg(p.bias + ncol + i) : 0.f;
          if (GROUPS == 2) {
            const float o0 = (a[i] - b[i]) + bs, o1 = (a[i] + b[i]) + bs;
            const float m = (p.relu_mode == 1 && !(o0 >= 0.f)) ? 0.f : 1.f;
            a[i] = p.relu_mode == 1 ? o0 * m : o0;
            b[i] = p.relu_mode == 1 ? o1 * m : o1;
          } else {
            a[i] = a[i] + bs;
            if (p.relu_mode == 2) a[i] = tanhf(a[i]);
          }
        }
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
          const float* o = g ? b : a;
          char* yp = p.y[g] + row * p.ldy_bytes[g] + int64_t(ncol) * ES;
          if constexpr (BF16) {
#pragma unroll
            for (int i = 0; i < CPW; i += 8) {
              float4 pk;
              pk.x = __uint_as_float(pack_bf16(o[i], o[i + 1]));
              pk.y = __uint_as_float(pack_bf16(o[i + 2], o[i + 3]));
              pk.z = __uint_as_float(pack_bf16(o[i + 4], o[i + 5]));
              pk.w = __uint_as_float(pack_bf16(o[i + 6], o[i + 7]));
              st_stream_v4(yp + i * 2, pk, pol_stream);
            }
          } else {
#pragma unroll
            for (int i = 0; i < CPW; i += 4)
              st_stream_v4(yp + i * 4, make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]), pol_stream);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS)
                 : "memory");
  }
}

static size_t smem_bytes_ws(int n_slabs, int n_tile, bool bf16, int stages) {
  return 1024 + size_t(stages) * STAGE_BYTES + size_t(n_slabs) * (bf16 ? 1 : 2) * n_tile * 128 +
         8 * (2 * WS_MAX_STAGES + 4) + 16;
}

template <int N_OUT, int GROUPS, bool BF16>
static int launch_ws(Params p, cudaStream_t st) {
  int stages = WS_MAX_STAGES;
  while (stages > 2 && smem_bytes_ws(p.n_slabs, N_OUT, BF16, stages) > 220 * 1024) --stages;
  p.stages = stages;
  const size_t smem = smem_bytes_ws(p.n_slabs, N_OUT, BF16, stages);
  auto kern = dense_tc_ws_kernel<N_OUT, GROUPS, BF16>;
  PGSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;
  const int n_col_tiles = (p.n_total + N_OUT - 1) / N_OUT;
  int64_t gx = sm_count() / n_col_tiles;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  kern<<<dim3(unsigned(gx), unsigned(n_col_tiles)), WS_THREADS, smem, st>>>(p);
  PGSD_LAUNCH_CHECK("dense_tc_ws_kernel");
  return PGSD_OK;
}

static size_t smem_bytes(int n_slabs, int n_tile, bool bf16) {
  return 1024 + STAGES * STAGE_BYTES + size_t(n_slabs) * (bf16 ? 1 : 2) * n_tile * 128 + 8 * (STAGES + 1) + 16;
}

template <int N_OUT, int GROUPS, bool BF16, int PFK = PF>
static int launch(const Params& p, cudaStream_t st) {
  const size_t smem = smem_bytes(p.n_slabs, N_OUT, BF16);
  auto kern = dense_tc_kernel<N_OUT, GROUPS, BF16, PFK>;
  PGSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;
  const int n_col_tiles = (p.n_total + N_OUT - 1) / N_OUT;
  int64_t gx = sm_count() / n_col_tiles;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  kern<<<dim3(unsigned(gx), unsigned(n_col_tiles)), THREADS, smem, st>>>(p);
  PGSD_LAUNCH_CHECK("dense_tc_kernel");
  return PGSD_OK;
}

static inline bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace tc

// Returns PGSD_OK and sets *handled = 1 when the tensor-core path ran; *handled = 0 means the
// problem is outside its envelope and the caller should use the FFMA kernel.
int dense_tc_try(const pgsd_dense_args* a, cudaStream_t st, int* handled) {
  using namespace tc;
  *handled = 0;
  const bool bf16 = a->dtype == PGSD_BF16;
  const int es = bf16 ? 2 : 4, ck = 128 / es, kal = 16 / es;
  const int n = a->n_out;
  // output tile width: the whole width when it is a supported tile, else 128-column tiles
  int n_tile = 0;
  if (n == 16 || n == 32 || n == 64 || n == 128) n_tile = n;
  else if (n > 128 && n % 128 == 0) n_tile = 128;
  if (n_tile == 0 || (bf16 && n_tile < 32)) return PGSD_OK;
  const int groups = a->combine ? 2 : 1;
  Params p{};
  p.n_rows = a->n_rows;
  p.n_total = n;
  p.relu_mode = a->relu_mode;
  p.bias = a->bias;
  for (int i = 0; i < groups; ++i) {
    if (!al16(a->y[i]) || ((a->ldy[i] * es) & 15)) return PGSD_OK;
    p.y[i] = static_cast<char*>(a->y[i]);
    p.ldy_bytes[i] = a->ldy[i] * es;
  }
  bool seen_group[2] = {false, false};
  for (int t = 0; t < a->n_terms; ++t) {
    const char* x = static_cast<const char*>(a->x[t]);
    if (!al16(x) || ((a->ldx[t] * es) & 15) || (a->k[t] % kal) || a->k[t] == 0) return PGSD_OK;
    for (int k0 = 0; k0 < a->k[t]; k0 += ck) {
      if (p.n_chunks >= MAX_CHUNKS) return PGSD_OK;
      const int kv = (a->k[t] - k0) < ck ? (a->k[t] - k0) : ck;
      const float* wb = a->w[t] + int64_t(k0) * a->ldw_k[t];
      int s = -1;
      for (int j = 0; j < p.n_slabs; ++j)
        if (p.w[j] == wb && p.ldw_k[j] == a->ldw_k[t] && p.ldw_n[j] == a->ldw_n[t] && p.wk[j] == kv) s = j;
      if (s < 0) {
        if (p.n_slabs >= MAX_SLABS) return PGSD_OK;
        s = p.n_slabs++;
        p.w[s] = wb, p.ldw_k[s] = a->ldw_k[t], p.ldw_n[s] = a->ldw_n[t], p.wk[s] = int8_t(kv);
      }
      const int c = p.n_chunks++;
      p.x[c] = x + int64_t(k0) * es;
      p.ldx_bytes[c] = a->ldx[t] * es;
      p.group[c] = int8_t(a->group[t]);
      p.slab[c] = int8_t(s);
      p.kvalid[c] = int8_t(kv);
      p.first[c] = seen_group[a->group[t]] ? 0 : 1;
      seen_group[a->group[t]] = true;
    }
  }
  if (groups == 2 && !(seen_group[0] && seen_group[1])) return PGSD_OK;
  if (smem_bytes(p.n_slabs, n_tile, bf16) > 220 * 1024) return PGSD_OK;
  int rc = PGSD_OK;
  // variant 4 selects the warp-specialised kernel (measured 0.589 ms vs 0.556 ms for the
  // synchronous one at 1M x (4 x 64) -> 64: both are instruction-issue bound on the hi/lo split and
  // address arithmetic, see profiles/README.md), so the synchronous kernel stays the default
  const bool sync_kernel = a->variant != 4;
  const bool deep = a->variant == 8;     // experiment: 8 chunks (128 KB per SM) of loads in flight instead of 4
#define PGSD_TC(N_)                                                                                   \
  if (deep) {                                                                                         \
    if (bf16) rc = groups == 2 ? launch<N_, 2, true, 8>(p, st) : launch<N_, 1, true, 8>(p, st);       \
    else rc = groups == 2 ? launch<N_, 2, false, 8>(p, st) : launch<N_, 1, false, 8>(p, st);          \
  } else if (sync_kernel) {                                                                           \
    if (bf16) rc = groups == 2 ? launch<N_, 2, true>(p, st) : launch<N_, 1, true>(p, st);             \
    else rc = groups == 2 ? launch<N_, 2, false>(p, st) : launch<N_, 1, false>(p, st);                \
  } else {                                                                                            \
    if (bf16) rc = groups == 2 ? launch_ws<N_, 2, true>(p, st) : launch_ws<N_, 1, true>(p, st);       \
    else rc = groups == 2 ? launch_ws<N_, 2, false>(p, st) : launch_ws<N_, 1, false>(p, st);          \
  }                                                                                                   \
  break;
  switch (n_tile) {
    case 16:
      if (sync_kernel) rc = groups == 2 ? launch<16, 2, false>(p, st) : launch<16, 1, false>(p, st);
      else rc = groups == 2 ? launch_ws<16, 2, false>(p, st) : launch_ws<16, 1, false>(p, st);
      break;
    case 32: PGSD_TC(32)
    case 64: PGSD_TC(64)
    default: PGSD_TC(128)
  }
#undef PGSD_TC
  if (rc == PGSD_OK) *handled = 1;
  return rc;
}

}  // namespace pgsd
