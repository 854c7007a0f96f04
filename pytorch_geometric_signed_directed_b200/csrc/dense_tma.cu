// TMA-fed, fully warp-specialised tcgen05 kernel of the dense feature transform (same math and
// envelope as dense_tc.cu: acc_g = sum_t X_t W_t, 3xTF32 for fp32 / kind::f16 for bf16, real/imag mix,
// bias, optional complex-ReLU mask).
//
// Why a second kernel: dense_tc_kernel moves the feature rows global -> registers -> shared memory with
// every thread playing every role, so the loads of a CTA advance in lock step with its per-chunk
// barrier, its single MMA-issuing thread and its epilogue (measured 0.52-0.56 ms on 1.54 GB = 2.9 TB/s,
// profiles/r01_sweep_dense_s14.jsonl).  Here the roles never wait for each other except through
// mbarriers, and the feature tiles never touch registers or L1 on their way in:
//   warp 0      : TMA producer -- one thread issues cp.async.bulk.tensor.2d per [128 rows x 128 B] chunk
//                 (SWIZZLE_128B tensor maps, one per term; rows / k past the end are zero-filled by
//                 the TMA unit) into an S-stage ring, completion on the stage's `full` mbarrier;
//   warps 10..17: converters (fp32 only) -- split the landed chunk into TF32 hi (in place, same bytes) and
//                 lo (a slot of a SHORT second ring: a lo image lives only from the split to the end of
//                 its MMAs, so 2-3 slots serve any number of landing stages and the shared memory
//                 they save buys landing stages = bytes in flight), fence.proxy.async, arrive on `conv`;
//                 bf16 chunks are already MMA operands: no converter warps at all;
//   warp 1      : MMA issuer -- waits `conv` (fp32) / `full` (bf16), issues the tcgen05.mma chain of the
//                 chunk, tcgen05.commit -> `empty` (stage back to the producer); the last chunk of a
//                 tile also commits to `acc_full`;
//   warps 2..9  : epilogue -- tcgen05.ld of the finished accumulators (TMEM double-buffered, so tile
//                 t+1 accumulates while tile t is stored; two warps per TMEM lane quarter), mix/bias/mask,
//                 swizzled staging tile in shared memory, one TMA store per [128 rows x 128 B] box
//                 (rows narrower than 128 B: per-lane streaming stores).
#include <cuda.h>

#include "tc_common.cuh"

namespace pgsd {
namespace tma {
using namespace tc;

constexpr int TILE_M = 128;
constexpr int MAX_TERMS = PGSD_DENSE_MAX_TERMS;
constexpr int MAX_CHUNKS = 2 * PGSD_DENSE_MAX_TERMS;
constexpr int MAX_SLABS = 16;
constexpr int MAX_STAGES = 12;
constexpr int MAX_LO = 6;
constexpr int CHUNK_BYTES = TILE_M * 128;
constexpr int EPI_WARPS = 8;    // two per TMEM lane quarter, alternating 16-column blocks
constexpr int CONV_WARPS = 8;

struct alignas(64) Params {
  CUtensorMap maps[MAX_TERMS];     // one [n_rows, k_t] tensor per term, box = [128 rows, 128 bytes]
  CUtensorMap out_maps[2];         // outputs, same box (epilogue: registers -> swizzled staging -> TMA store)
  int64_t n_rows;
  int32_t n_chunks, n_slabs, relu_mode, n_total, stages, lo_stages;
  int16_t k0[MAX_CHUNKS];          // first k (elements) of the chunk inside its term
  int8_t term[MAX_CHUNKS];
  int8_t group[MAX_CHUNKS];
  int8_t slab[MAX_CHUNKS];
  int8_t first[MAX_CHUNKS];        // first chunk of its group inside a tile -> overwrite TMEM
  const float* w[MAX_SLABS];       // weight base + k0 * ldw_k
  int64_t ldw_k[MAX_SLABS], ldw_n[MAX_SLABS];
  int8_t wk[MAX_SLABS];            // valid k rows of the slab
  const float* bias;
  char* y[2];
  int64_t ldy_bytes[2];
  int32_t conv_groups;             // converter warps work on this many chunks at once (1, 2 or 4)
  int32_t epi_rb, epi_ns;          // epilogue staging: boxes per group and round, staging sets (1 or 2)
  int32_t dbg;                     // timing experiments only (variant bits 16-18): 1 no stores, 2 no split, 4 no MMAs
};

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, int c0, int c1, uint32_t src) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(
                   reinterpret_cast<uint64_t>(map)),
               "r"(c0), "r"(c1), "r"(src)
               : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the staging buffers of all committed stores have been read (they may be overwritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() {
  asm volatile("bar.sync 1, %0;" ::"n"(8 * 32) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
// waits that span a whole pipeline fill (and much more under a profiler): bounded, with back-off
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try(bar, parity)) {
    if (++spins > 64) __nanosleep(32);
    if (spins > (1u << 26)) __trap();
  }
}

// TANH: the tanh epilogue (SGCN.py:93-96) is a compile-time variant -- as a run-time branch its inlined
// tanhf cost the epilogue-bound bf16 instantiation 20 % (0.19 -> 0.235 ms, session 32)
// WIDE (fp32 only): [W_hi | W_lo] is ONE B operand of 2 * N_OUT rows, so hi * W_hi and hi * W_lo are a single MMA into
// two TMEM column ranges (summed in the epilogue) and the landed chunk is read by the tensor core ONCE as `hi` --
// RAW, without the in-place TF32 rounding: kind::tf32 ignores the 13 low mantissa bits, i.e. hi = trunc19(x), and the
// converters only produce lo = tf32(x - trunc19(x)) (one shared-memory write per chunk instead of two).  The kernel is
// shared-memory-bandwidth bound (ncu: LSU-shared 48 %, ~1.2 MB of shared-memory traffic per 128-row tile of a MagNet
// layer); this removes a fifth of it.  Error budget per product: (lo - tf32(lo)) w_hi <= 2^-21, hi (w_lo - tf32(w_lo))
// <= 2^-22, dropped lo * w_lo <= 2^-21 -- fp32-class, checked against fp64 at 2e-6 like the other paths.
template <int N_OUT, int GROUPS, bool BF16, bool TANH, bool WIDE>
__global__ void __launch_bounds__((2 + EPI_WARPS + (BF16 ? 0 : CONV_WARPS)) * 32, 1)
    dense_tma_kernel(const __grid_constant__ Params p) {
  constexpr int NCV = BF16 ? 0 : CONV_WARPS;
  constexpr int THREADS = (2 + EPI_WARPS + NCV) * 32;
  constexpr int ES = BF16 ? 2 : 4;
  constexpr int EPC = 16 / ES;
  constexpr int CK = 128 / ES;
  constexpr int STAGE_BYTES = CHUNK_BYTES;       // landing ring; fp32: becomes the hi image in place
  constexpr int HALF = N_OUT * 128;
  constexpr int SLAB_BYTES = BF16 ? HALF : 2 * HALF;
  constexpr int UMMA_K_BYTES = 32;
  static_assert(!WIDE || !BF16, "WIDE is the fp32 (3xTF32) path");
  constexpr int GCOLS = WIDE ? 2 * N_OUT : N_OUT;   // TMEM columns per group
  constexpr int ACC_COLS = GROUPS * GCOLS;
  // output rows of at least 128 B leave through TMA stores; narrower ones through per-lane stores
  constexpr bool TMA_STORE = N_OUT * ES >= 128;
  constexpr int CBR = 128 / (16 * ES);           // 16-column blocks per 128-byte box row: 2 (fp32) / 4 (bf16)
  constexpr uint32_t TMEM_COLS = (2 * ACC_COLS <= 32) ? 32 : (2 * ACC_COLS <= 64) ? 64
                               : (2 * ACC_COLS <= 128) ? 128 : (2 * ACC_COLS <= 256) ? 256 : 512;
  constexpr uint32_t FMT = BF16 ? 1u : 2u;
  constexpr uint32_t IDESC = (1u << 4) | (FMT << 7) | (FMT << 10) | (uint32_t(N_OUT >> 3) << 17) |
                             (uint32_t(TILE_M >> 4) << 24);
  constexpr uint32_t IDESC_WIDE = (1u << 4) | (FMT << 7) | (FMT << 10) | (uint32_t((2 * N_OUT) >> 3) << 17) |
                                  (uint32_t(TILE_M >> 4) << 24);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int S = p.stages, L = BF16 ? 0 : p.lo_stages;
  uint8_t* a_stage = smem;
  uint8_t* lo_ring = smem + S * STAGE_BYTES;
  uint8_t* out_stage = lo_ring + L * CHUNK_BYTES;            // GROUPS boxes of [128 rows x 128 B]
  uint8_t* w_smem = out_stage + (TMA_STORE ? p.epi_ns * p.epi_rb * GROUPS * CHUNK_BYTES : 0);
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + p.n_slabs * SLAB_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * MAX_STAGES + MAX_LO + 4);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n0 = blockIdx.y * N_OUT;
  const uint32_t a_addr = smem_u32(a_stage), lo_addr = smem_u32(lo_ring), w_addr = smem_u32(w_smem);
  const uint32_t bar_full = smem_u32(bars), bar_conv = bar_full + 8 * MAX_STAGES, bar_empty = bar_conv + 8 * MAX_STAGES;
  const uint32_t bar_lo_empty = bar_empty + 8 * MAX_STAGES;
  const uint32_t bar_acc_full = bar_lo_empty + 8 * MAX_LO, bar_acc_empty = bar_acc_full + 16;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (tid == 32) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_conv + 8 * s, NCV > 0 ? NCV / p.conv_groups : 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int s = 0; s < MAX_LO; ++s) mbar_init(bar_lo_empty + 8 * s, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_acc_full + 8 * a, 1);
      mbar_init(bar_acc_empty + 8 * a, EPI_WARPS);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // weights -> K-major SW128 images (element (n, k) in 16-byte unit (k / EPC) ^ (n & 7) of row n)
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
  const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;

  if (warp == 0) {
    // ======================================================================== TMA producer
    uint32_t uses = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int c = 0; c < p.n_chunks; ++c, ++uses) {
        const uint32_t stage = uses % S;
        if (lane == 0) {
          if (uses >= uint32_t(S)) mbar_wait_relaxed(bar_empty + 8 * stage, ((uses / S) - 1) & 1);
          mbar_expect_tx(bar_full + 8 * stage, CHUNK_BYTES);
          tma_load_2d(a_addr + stage * STAGE_BYTES, &p.maps[p.term[c]], p.k0[c], int(tile * TILE_M),
                      bar_full + 8 * stage);
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ========================================================================== MMA issuer
    // The whole warp runs this loop converged (all operands stay warp-uniform); one elected lane issues.
    const uint32_t a_lo32 = desc_lo(a_addr), l_lo32 = BF16 ? 0u : desc_lo(lo_addr), w_lo32 = desc_lo(w_addr);
    constexpr uint32_t KSTEP = UMMA_K_BYTES >> 4;              // descriptor units per MMA along K
    uint32_t uses = 0, it = 0, stage = 0, ls = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      if (it >= 2) mbar_wait_relaxed(bar_acc_empty + 8 * acc, ((it >> 1) - 1) & 1);   // epilogue drained this buffer
      for (int c = 0; c < p.n_chunks; ++c, ++uses) {
        const uint32_t d = tmem_base + acc * ACC_COLS + uint32_t(p.group[c]) * GCOLS;
        if constexpr (WIDE) {
          // hi MMAs as soon as the chunk has LANDED (they read the raw words), lo MMAs once the converters are done
          const uint32_t da_hi = a_lo32 + stage * (STAGE_BYTES >> 4);
          const uint32_t da_lo = l_lo32 + ls * (CHUNK_BYTES >> 4);
          const uint32_t dw = w_lo32 + uint32_t(p.slab[c]) * (SLAB_BYTES >> 4);
          const uint32_t acc0 = p.first[c] ? 0u : 1u;
          mbar_wait_relaxed(bar_full + 8 * stage, (uses / S) & 1);
          tc_fence_after();
          if (elect_one() && !(p.dbg & 4)) {
#pragma unroll
            for (int j = 0; j < 128 / UMMA_K_BYTES; ++j)
              mma_lo<false>(d, da_hi + j * KSTEP, dw + j * KSTEP, IDESC_WIDE, j == 0 ? acc0 : 1u);
          }
          __syncwarp();
          mbar_wait_relaxed(bar_conv + 8 * stage, (uses / S) & 1);
          tc_fence_after();
          if (elect_one()) {
            if (!(p.dbg & 4)) {
#pragma unroll
              for (int j = 0; j < 128 / UMMA_K_BYTES; ++j)
                mma_lo<false>(d, da_lo + j * KSTEP, dw + j * KSTEP, IDESC, 1u);
            }
            tc_commit(bar_empty + 8 * stage);
            tc_commit(bar_lo_empty + 8 * ls);
            if (c == p.n_chunks - 1) tc_commit(bar_acc_full + 8 * acc);
          }
          __syncwarp();
          stage = (stage + 1 == uint32_t(S)) ? 0u : stage + 1;
          ls = (ls + 1 == uint32_t(L)) ? 0u : ls + 1;
          continue;
        }
        mbar_wait_relaxed((BF16 ? bar_full : bar_conv) + 8 * stage, (uses / S) & 1);
        tc_fence_after();
        const uint32_t da_hi = a_lo32 + stage * (STAGE_BYTES >> 4);
        const uint32_t da_lo = l_lo32 + ls * (CHUNK_BYTES >> 4);
        const uint32_t dw_hi = w_lo32 + uint32_t(p.slab[c]) * (SLAB_BYTES >> 4), dw_lo = dw_hi + (HALF >> 4);
        const uint32_t acc0 = p.first[c] ? 0u : 1u;
        if (elect_one()) {
          if (!(p.dbg & 4)) {
#pragma unroll
            for (int j = 0; j < 128 / UMMA_K_BYTES; ++j) {
              if constexpr (BF16) {
                mma_lo<true>(d, da_hi + j * KSTEP, dw_hi + j * KSTEP, IDESC, j == 0 ? acc0 : 1u);
              } else {
                mma_lo<false>(d, da_hi + j * KSTEP, dw_hi + j * KSTEP, IDESC, j == 0 ? acc0 : 1u);
                mma_lo<false>(d, da_lo + j * KSTEP, dw_hi + j * KSTEP, IDESC, 1u);
                mma_lo<false>(d, da_hi + j * KSTEP, dw_lo + j * KSTEP, IDESC, 1u);
              }
            }
          }
          tc_commit(bar_empty + 8 * stage);
          if constexpr (!BF16) tc_commit(bar_lo_empty + 8 * ls);
          if (c == p.n_chunks - 1) tc_commit(bar_acc_full + 8 * acc);
        }
        __syncwarp();
        stage = (stage + 1 == uint32_t(S)) ? 0u : stage + 1;
        if constexpr (!BF16) ls = (ls + 1 == uint32_t(L)) ? 0u : ls + 1;
      }
    }
  } else if (warp < 2 + EPI_WARPS) {
    // ============================================================================ epilogue
    const int q = warp & 3;                      // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;            // which of the two warps of the quarter: even / odd column blocks
    constexpr int NCB = N_OUT / 16;
    const uint64_t pol_stream = policy_evict_first();
    uint32_t it = 0, rounds = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const uint32_t acc = it & 1;
      mbar_wait_relaxed(bar_acc_full + 8 * acc, (it >> 1) & 1);
      tc_fence_after();
      const int64_t row = tile * TILE_M + q * 32 + lane;
      const uint32_t taddr = tmem_base + (uint32_t(q * 32) << 16) + acc * ACC_COLS;
      if constexpr (TMA_STORE) {
        // rounds of RB 128-byte-wide boxes per group: tcgen05.ld -> mix/bias/mask -> swizzled staging
        // (conflict-free: lane = row, 16-byte unit ^ (row & 7)) -> one TMA store per box.  Full-line,
        // asynchronous writes instead of 32 scattered 16-byte pieces per store instruction.  With two
        // staging sets the stores of round r are still reading set r & 1 while round r + 1 fills the other.
        constexpr int NB = NCB / CBR;                          // boxes per group and tile
        constexpr int NCI = (GROUPS == 1 && BF16) ? 2 : 1;     // column blocks per TMEM wait
        const int RB = p.epi_rb, NS = p.epi_ns;
        const uint32_t set_bytes = uint32_t(GROUPS * RB) * CHUNK_BYTES;
        const uint32_t row_off = uint32_t(q * 32 + lane) * 128;
#pragma unroll 1
        for (int rd = 0; rd < NB / RB; ++rd, ++rounds) {
          const uint32_t set_addr = smem_u32(out_stage) + (NS == 2 ? (rounds & 1) * set_bytes : 0u);
          if (warp == 2 && lane == 0) {                        // this set's previous stores have read it
            if (NS == 2) tma_store_wait_read1(); else tma_store_wait_read();
          }
          epi_bar_sync();
#pragma unroll 1
          for (int kk0 = half * NCI; kk0 < RB * CBR; kk0 += 2 * NCI) {
            float a[NCI][16], b[NCI][16], bs[NCI][16];
#pragma unroll
            for (int n = 0; n < NCI; ++n) {
              const int cb = rd * RB * CBR + kk0 + n;
              tmem_ld<16>(taddr + cb * 16, a[n]);
              if (GROUPS == 2) tmem_ld<16>(taddr + GCOLS + cb * 16, b[n]);
            }
            if constexpr (WIDE) {
              // second column range of every group: hi * W_lo (NCI == 1 on the fp32 path)
              float a2[16], b2[16];
              const int cb = rd * RB * CBR + kk0;
              tmem_ld<16>(taddr + N_OUT + cb * 16, a2);
              if (GROUPS == 2) tmem_ld<16>(taddr + GCOLS + N_OUT + cb * 16, b2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                a[0][i] += a2[i];
                if (GROUPS == 2) b[0][i] += b2[i];
              }
            }
#pragma unroll
            for (int n = 0; n < NCI; ++n)
#pragma unroll
              for (int i = 0; i < 16; ++i) {       // issued while the TMEM loads are in flight
                const int col = n0 + (rd * RB * CBR + kk0 + n) * 16 + i;
                bs[n][i] = (p.bias && col < p.n_total) ? __ldg(p.bias + col) : 0.f;
              }
            tmem_ld_wait();
            if (rd == NB / RB - 1 && kk0 + 2 * NCI >= RB * CBR) {   // last TMEM read of this warp for this buffer
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
            }
#pragma unroll
            for (int n = 0; n < NCI; ++n) {
              const int kk = kk0 + n, bx = kk / CBR, k = kk % CBR;
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                if (GROUPS == 2) {
                  const float o0 = (a[n][i] - b[n][i]) + bs[n][i], o1 = (a[n][i] + b[n][i]) + bs[n][i];
                  const float m = (p.relu_mode == 1 && !(o0 >= 0.f)) ? 0.f : 1.f;
                  a[n][i] = p.relu_mode == 1 ? o0 * m : o0;
                  b[n][i] = p.relu_mode == 1 ? o1 * m : o1;
                } else {
                  a[n][i] = a[n][i] + bs[n][i];
                }
              }
              if constexpr (TANH) {
#pragma unroll
                for (int i = 0; i < 16; ++i) a[n][i] = tanhf(a[n][i]);
              }
#pragma unroll
              for (int g = 0; g < GROUPS; ++g) {
                const float* o = g ? b[n] : a[n];
                const uint32_t base = set_addr + uint32_t(g * RB + bx) * CHUNK_BYTES + row_off;
                if constexpr (BF16) {
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    float4 pk;
                    pk.x = __uint_as_float(pack_bf16(o[8 * j], o[8 * j + 1]));
                    pk.y = __uint_as_float(pack_bf16(o[8 * j + 2], o[8 * j + 3]));
                    pk.z = __uint_as_float(pack_bf16(o[8 * j + 4], o[8 * j + 5]));
                    pk.w = __uint_as_float(pack_bf16(o[8 * j + 6], o[8 * j + 7]));
                    sts_v4(base + (uint32_t((k * 2 + j) ^ (lane & 7)) << 4), pk);
                  }
                } else {
#pragma unroll
                  for (int j = 0; j < 4; ++j)
                    sts_v4(base + (uint32_t((k * 4 + j) ^ (lane & 7)) << 4),
                           make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]));
                }
              }
            }
          }
          fence_async_smem();
          epi_bar_sync();
          if (warp == 2 && lane == 0 && !(p.dbg & 1)) {
            for (int g = 0; g < GROUPS; ++g)
              for (int bx = 0; bx < RB; ++bx)
                tma_store_2d(&p.out_maps[g], n0 + (rd * RB + bx) * CBR * 16, int(tile * TILE_M),
                             set_addr + uint32_t(g * RB + bx) * CHUNK_BYTES);
            tma_store_commit();
          }
        }
      } else {
      if (half >= NCB) {                         // N_OUT = 16: the odd warps have no column block
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
      }
#pragma unroll 1
      for (int cb = half; cb < NCB; cb += 2) {
        float a[16], b[16], bs[16];
        tmem_ld<16>(taddr + cb * 16, a);
        if (GROUPS == 2) tmem_ld<16>(taddr + N_OUT + cb * 16, b);
#pragma unroll
        for (int i = 0; i < 16; ++i)             // issued while the TMEM loads are in flight
          bs[i] = (p.bias && n0 + cb * 16 + i < p.n_total) ? __ldg(p.bias + n0 + cb * 16 + i) : 0.f;
        tmem_ld_wait();
        if (cb + 2 >= NCB) {                     // everything this warp reads of the buffer is in registers
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(bar_acc_empty + 8 * acc);
        }
        const int ncol = n0 + cb * 16;
        if (row < p.n_rows && ncol < p.n_total && !(p.dbg & 1)) {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            if (GROUPS == 2) {
              const float o0 = (a[i] - b[i]) + bs[i], o1 = (a[i] + b[i]) + bs[i];
              const float m = (p.relu_mode == 1 && !(o0 >= 0.f)) ? 0.f : 1.f;
              a[i] = p.relu_mode == 1 ? o0 * m : o0;
              b[i] = p.relu_mode == 1 ? o1 * m : o1;
            } else {
              a[i] = a[i] + bs[i];
            }
          }
          if constexpr (TANH) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = tanhf(a[i]);
          }
#pragma unroll
          for (int g = 0; g < GROUPS; ++g) {
            const float* o = g ? b : a;
            char* yp = p.y[g] + row * p.ldy_bytes[g] + int64_t(ncol) * ES;
            if constexpr (BF16) {
#pragma unroll
              for (int i = 0; i < 16; i += 8) {
                float4 pk;
                pk.x = __uint_as_float(pack_bf16(o[i], o[i + 1]));
                pk.y = __uint_as_float(pack_bf16(o[i + 2], o[i + 3]));
                pk.z = __uint_as_float(pack_bf16(o[i + 4], o[i + 5]));
                pk.w = __uint_as_float(pack_bf16(o[i + 6], o[i + 7]));
                st_stream_v4(yp + i * 2, pk, pol_stream);
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; i += 4)
                st_stream_v4(yp + i * 4, make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]), pol_stream);
            }
          }
        }
      }
      }
    }
    if (TMA_STORE && warp == 2 && lane == 0) tma_store_wait_all();   // global writes complete before exit
  } else {
    // ================================================================ converters (fp32 only)
    if constexpr (!BF16) {
      // G groups of NCV / G warps; group g splits the chunks u = g (mod G), so G chunks are in
      // conversion at once and the LDS -> split -> STS -> fence.proxy.async -> arrive latency of one
      // chunk overlaps the next ones' (with one group that chain paced the whole kernel)
      const int cw = warp - (2 + EPI_WARPS);                // 0 .. NCV-1
      const int G = p.conv_groups, wpg = NCV / G;
      const int grp = cw % G, tig = (cw / G) * 32 + lane;   // thread inside its group
      const uint32_t gstride = uint32_t(wpg) * 32 * 16;     // bytes between the units of one thread
      uint32_t uses = 0, stage = 0, ls = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int c = 0; c < p.n_chunks; ++c, ++uses) {
          if (int(uses % uint32_t(G)) == grp) {
            mbar_wait_relaxed(bar_full + 8 * stage, (uses / S) & 1);
            // the lo slot is free once the MMAs of the chunk that used it last (L chunks ago) completed
            if (uses >= uint32_t(L)) mbar_wait_relaxed(bar_lo_empty + 8 * ls, ((uses / L) - 1) & 1);
            const uint32_t hi = a_addr + stage * STAGE_BYTES + tig * 16, lo = lo_addr + ls * CHUNK_BYTES + tig * 16;
            if (!(p.dbg & 2)) {
              for (int b = 0; b < G; ++b) {                 // 4 G units per thread, 4 at a time
                const uint32_t ob = uint32_t(b) * 4 * gstride;
                float4 v[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = lds_v4(hi + ob + i * gstride);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                  // the split is position-independent: unit q of the landed (swizzled) image stays unit q
                  float4 vh, vl;
                  if constexpr (WIDE) {
                    // hi = what the tensor core sees of the raw word (13 low mantissa bits ignored): nothing to store
                    auto trunc19 = [](float f) { return __uint_as_float(__float_as_uint(f) & 0xffffe000u); };
                    vl.x = to_tf32(v[i].x - trunc19(v[i].x)), vl.y = to_tf32(v[i].y - trunc19(v[i].y));
                    vl.z = to_tf32(v[i].z - trunc19(v[i].z)), vl.w = to_tf32(v[i].w - trunc19(v[i].w));
                  } else {
                    vh.x = to_tf32(v[i].x), vh.y = to_tf32(v[i].y), vh.z = to_tf32(v[i].z), vh.w = to_tf32(v[i].w);
                    vl.x = to_tf32(v[i].x - vh.x), vl.y = to_tf32(v[i].y - vh.y);
                    vl.z = to_tf32(v[i].z - vh.z), vl.w = to_tf32(v[i].w - vh.w);
                    sts_v4(hi + ob + i * gstride, vh);
                  }
                  sts_v4(lo + ob + i * gstride, vl);
                }
              }
            }
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_conv + 8 * stage);
          }
          stage = (stage + 1 == uint32_t(S)) ? 0u : stage + 1;
          ls = (ls + 1 == uint32_t(L)) ? 0u : ls + 1;
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

static size_t smem_bytes(int n_slabs, int n_tile, bool bf16, int stages, int lo_stages, int out_boxes) {
  return 1024 + size_t(stages + (bf16 ? 0 : lo_stages) + out_boxes) * CHUNK_BYTES +
         size_t(n_slabs) * (bf16 ? 1 : 2) * n_tile * 128 + 8 * (3 * MAX_STAGES + MAX_LO + 4) + 16;
}

template <int N_OUT, int GROUPS, bool BF16, bool TANH = false, bool WIDE = false>
static int launch(Params& p, cudaStream_t st, int want_stages, int want_lo, int want_cg, int want_epi) {
  // experiment knobs ride in the variant word (bits 8-11: lo slots, bits 12-15: landing stages,
  // bits 20-22: converter groups); 0 = default
  // Ring invariants (mbarrier waits only tell odd from even phases, so no waiter may run a whole phase
  // ahead): converter groups <= landing stages <= 2 * lo slots.  Preferred: 2 groups, 3 lo slots; shapes
  // whose weights leave little shared memory fall back to 1 group, 2 lo slots.
  int cg = (want_cg == 1 || want_cg == 2 || want_cg == 4) ? want_cg : 2;
  int lo = BF16 ? 0 : (want_lo >= 1 && want_lo <= MAX_LO ? want_lo : cg + 1);
  const size_t budget = 220 * 1024;
  constexpr bool TMA_STORE = N_OUT * (BF16 ? 2 : 4) >= 128;
  constexpr int NB = TMA_STORE ? N_OUT * (BF16 ? 2 : 4) / 128 : 0;    // 128-byte boxes per group and tile
  // epilogue staging, best first: the whole tile per round in two sets, ... , one box per round in one set;
  // a candidate must leave room for 4 landing stages (3 when the weights are large)
  int rb = 1, ns = 1;
  if (TMA_STORE) {
    const int cand[4][2] = {{NB, 2}, {NB, 1}, {1, 2}, {1, 1}};
    bool found = false;
    for (int min_stages = 4; min_stages >= 3 && !found; --min_stages)
      for (int c = 0; c < 4 && !found; ++c)
        if (smem_bytes(p.n_slabs, N_OUT, BF16, min_stages, lo, cand[c][0] * cand[c][1] * GROUPS) <= budget) {
          rb = cand[c][0], ns = cand[c][1];
          found = true;
        }
  }
  if (want_epi == 1) rb = 1, ns = 1;                     // experiment knob (variant bits 24-25)
  if (want_epi == 2) rb = 1, ns = 2;
  if (want_epi == 3) rb = NB > 0 ? NB : 1, ns = 1;
  int out_boxes = TMA_STORE ? rb * ns * GROUPS : 0;
  if (!BF16 && smem_bytes(p.n_slabs, N_OUT, BF16, 3, lo, out_boxes) > budget) {
    cg = 1, lo = 2, rb = 1, ns = 1;
    out_boxes = TMA_STORE ? GROUPS : 0;
  }
  int stages = (want_stages >= 2 && want_stages <= MAX_STAGES) ? want_stages : MAX_STAGES;
  if (!BF16 && stages > 2 * lo) stages = 2 * lo;
  while (stages > 2 && smem_bytes(p.n_slabs, N_OUT, BF16, stages, lo, out_boxes) > budget) --stages;
  if (smem_bytes(p.n_slabs, N_OUT, BF16, stages, lo, out_boxes) > budget) return -1;
  if (cg > stages) cg = stages >= 2 ? 2 : 1;
  p.conv_groups = cg;
  p.stages = stages;
  p.lo_stages = lo;
  p.epi_rb = rb, p.epi_ns = ns;
  const size_t smem = smem_bytes(p.n_slabs, N_OUT, BF16, stages, lo, out_boxes);
  auto kern = dense_tma_kernel<N_OUT, GROUPS, BF16, TANH, WIDE>;
  PGSD_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
  const int64_t n_tiles = (p.n_rows + TILE_M - 1) / TILE_M;
  const int n_col_tiles = (p.n_total + N_OUT - 1) / N_OUT;
  int64_t gx = sm_count() / n_col_tiles;
  if (gx < 1) gx = 1;
  if (gx > n_tiles) gx = n_tiles;
  constexpr int THREADS = (2 + EPI_WARPS + (BF16 ? 0 : CONV_WARPS)) * 32;
  kern<<<dim3(unsigned(gx), unsigned(n_col_tiles)), THREADS, smem, st>>>(p);
  PGSD_LAUNCH_CHECK("dense_tma_kernel");
  return PGSD_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point query: no -lcuda at link time
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

static inline bool al16(const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; }

}  // namespace tma

// Same contract as dense_tc_try: *handled = 1 when this kernel ran.
int dense_tma_try(const pgsd_dense_args* a, cudaStream_t st, int* handled) {
  using namespace tma;
  *handled = 0;
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return PGSD_OK;
  const bool bf16 = a->dtype == PGSD_BF16;
  const int es = bf16 ? 2 : 4, ck = 128 / es, kal = 16 / es;
  const int n = a->n_out;
  int n_tile = 0;
  if (n == 16 || n == 32 || n == 64 || n == 128) n_tile = n;
  else if (n > 128 && n % 128 == 0) n_tile = 128;
  if (n_tile == 0 || (bf16 && n_tile < 32)) return PGSD_OK;
  if (a->n_rows <= 0 || a->n_rows >= (int64_t(1) << 31)) return PGSD_OK;
  const int groups = a->combine ? 2 : 1;
  static thread_local Params p;                 // 64-byte aligned, ~2.7 KB: keep it off the stack
  p = Params{};
  p.n_rows = a->n_rows;
  p.n_total = n;
  p.relu_mode = a->relu_mode;
  p.bias = a->bias;
  for (int i = 0; i < groups; ++i) {
    if (!al16(a->y[i]) || ((a->ldy[i] * es) & 15)) return PGSD_OK;
    p.y[i] = static_cast<char*>(a->y[i]);
    p.ldy_bytes[i] = a->ldy[i] * es;
    if (n_tile * es >= 128) {                 // TMA-store epilogue: [n_rows, n_out] tensor, box = [128 rows, 128 B]
      const cuuint64_t gdim[2] = {cuuint64_t(n), cuuint64_t(a->n_rows)};
      const cuuint64_t gstride[1] = {cuuint64_t(a->ldy[i]) * es};
      const cuuint32_t box[2] = {cuuint32_t(ck), cuuint32_t(TILE_M)};
      const cuuint32_t estride[2] = {1, 1};
      const CUresult r = enc(&p.out_maps[i], bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32,
                             2, a->y[i], gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return PGSD_OK;
    }
  }
  bool seen_group[2] = {false, false};
  for (int t = 0; t < a->n_terms; ++t) {
    const char* x = static_cast<const char*>(a->x[t]);
    if (!al16(x) || ((a->ldx[t] * es) & 15) || (a->k[t] % kal) || a->k[t] == 0 || a->k[t] > 32000) return PGSD_OK;
    const cuuint64_t gdim[2] = {cuuint64_t(a->k[t]), cuuint64_t(a->n_rows)};
    const cuuint64_t gstride[1] = {cuuint64_t(a->ldx[t]) * es};
    const cuuint32_t box[2] = {cuuint32_t(ck), cuuint32_t(TILE_M)};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = enc(&p.maps[t], bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2,
                           const_cast<char*>(x), gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return PGSD_OK;     // e.g. a stride the TMA unit cannot express -> other kernels
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
      p.term[c] = int8_t(t);
      p.k0[c] = int16_t(k0);
      p.group[c] = int8_t(a->group[t]);
      p.slab[c] = int8_t(s);
      p.first[c] = seen_group[a->group[t]] ? 0 : 1;
      seen_group[a->group[t]] = true;
    }
  }
  if (groups == 2 && !(seen_group[0] && seen_group[1])) return PGSD_OK;
  int rc = -1;
  const int ws = (a->variant >> 12) & 0xf, wl = (a->variant >> 8) & 0xf;
  p.dbg = (a->variant >> 16) & 7;
  const int wg = (a->variant >> 20) & 7, we = (a->variant >> 24) & 3;
  const bool tanh_epi = a->relu_mode == 2;
  if (tanh_epi && (bf16 || groups == 2)) return PGSD_OK;    // other kernels handle it
  // variant bit 26: the legacy split (hi rounded in place, three MMAs per k-step) for A/B timing
  const bool wide_ok = !bf16 && ((a->variant >> 26) & 1) == 0;
#define PGSD_TMA(N_, WG2_, WG1_)                                                                             \
  if (bf16) rc = groups == 2 ? launch<N_, 2, true>(p, st, ws, wl, wg, we) : launch<N_, 1, true>(p, st, ws, wl, wg, we); \
  else if (groups == 2) rc = (wide_ok && WG2_) ? launch<N_, 2, false, false, WG2_>(p, st, ws, wl, wg, we)     \
                                               : launch<N_, 2, false>(p, st, ws, wl, wg, we);                \
  else if (tanh_epi) rc = (wide_ok && WG1_) ? launch<N_, 1, false, true, WG1_>(p, st, ws, wl, wg, we)        \
                                            : launch<N_, 1, false, true>(p, st, ws, wl, wg, we);             \
  else rc = (wide_ok && WG1_) ? launch<N_, 1, false, false, WG1_>(p, st, ws, wl, wg, we)                     \
                              : launch<N_, 1, false>(p, st, ws, wl, wg, we);                                 \
  break;
  switch (n_tile) {
    case 16:
      if (groups == 2) rc = launch<16, 2, false>(p, st, ws, wl, wg, we);
      else rc = tanh_epi ? launch<16, 1, false, true>(p, st, ws, wl, wg, we) : launch<16, 1, false>(p, st, ws, wl, wg, we);
      break;
    // WIDE needs 2 buffers x groups x 2 x N_OUT TMEM columns <= 512
    case 32: PGSD_TMA(32, true, true)
    case 64: PGSD_TMA(64, true, true)
    default: PGSD_TMA(128, false, true)
  }
#undef PGSD_TMA
  if (rc == -1) return PGSD_OK;                  // does not fit in shared memory
  if (rc == PGSD_OK) *handled = 1;
  return rc;
}

}  // namespace pgsd
