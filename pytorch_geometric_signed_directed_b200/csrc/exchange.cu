// Shard push: the all-gather of the node-range sharded path as one kernel over NVLink peer memory.
// (the reference is single-device, so there is no reference counterpart.)
//
// Why a kernel and not cudaMemcpyPeer / NCCL: the aggregation that runs beside the exchange is
// HBM-bound, so the exchange is paid for in HBM traffic as much as in NVLink time.  A copy-engine or
// send/recv all-gather reads the local shard world-1 times (once per peer); here every 16-byte piece
// is loaded ONCE and stored to all peers from registers (or once to an NVSwitch multicast address),
// the receive buffers are written in place (no pack / unpack pass), and completion is signalled per
// row slice so the consumer can start on the first slice while the rest is in flight.
#include "tc_common.cuh"

namespace pgsd {

struct PushParams {
  int world, rank, n_tensors, n_slices, include_self;
  int cpr;        // 16-byte pieces per row
  int cpr_shift;  // log2(cpr) when cpr is a power of two, else -1
  uint32_t seq;
  const char* src[2];
  int64_t ld_src[2];
  char* dst[2][PGSD_MAX_RANKS];
  int64_t ld_dst[2];
  char* mc_dst[2];
  int64_t slice_row[PGSD_MAX_SLICES + 1];
  uint32_t* flag[PGSD_MAX_RANKS];
  uint32_t* counters;
  uint32_t* started;   // local word set to seq once every CTA of the push is resident (NULL = off)
};

__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v);
// Every CTA checks in when it starts; the last one publishes `started`.  The consumer stream waits for it before
// it launches the (persistent, SM-filling) aggregation, so the push CTAs always get their SM slots first.
__device__ __forceinline__ void announce_start(const PushParams& p) {
  if (p.started == nullptr || threadIdx.x != 0) return;
  const unsigned old = atomicAdd(p.counters + PGSD_MAX_SLICES, 1u);
  if (old == gridDim.x - 1) {
    p.counters[PGSD_MAX_SLICES] = 0;
    st_release_sys_u32(p.started, p.seq);
  }
}

__device__ __forceinline__ float4 ld_once_v4(const void* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_peer_v4(void* p, float4 v) {
  asm volatile("st.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_multicast_v4(void* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z),
               "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void st_release_sys_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// U pieces in flight per thread; MC = multicast stores; CONTIG = rows are contiguous on both sides (offset = piece
// index * 16, no row / column split and no per-piece offset registers).  256 threads x <= 80 registers: one CTA
// fits into the register space ONE aggregation CTA (256 threads x 80 registers) leaves free, which is what
// pgsd_spmm_args.grid_reserve counts.
template <int U, bool MC, bool CONTIG>
__global__ void __launch_bounds__(256) shard_push_kernel(const __grid_constant__ PushParams p) {
  announce_start(p);
  const int64_t nthreads = int64_t(gridDim.x) * blockDim.x;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n_peers = p.include_self ? p.world : p.world - 1;
  const int first = p.include_self ? 0 : 1;

  for (int s = 0; s < p.n_slices; ++s) {
    const int64_t r0 = p.slice_row[s];
    const int64_t total = (p.slice_row[s + 1] - r0) * p.cpr;
    for (int t = 0; t < p.n_tensors; ++t) {
      const char* src = p.src[t] + r0 * p.ld_src[t];
      const int64_t dst_off0 = r0 * p.ld_dst[t];
      for (int64_t i = tid; i < total; i += nthreads * U) {
        float4 v[U];
        if constexpr (CONTIG) {
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t idx = i + int64_t(u) * nthreads;
            if (idx < total) v[u] = ld_once_v4(src + idx * 16);
          }
          if constexpr (MC) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
              const int64_t idx = i + int64_t(u) * nthreads;
              if (idx < total) st_multicast_v4(p.mc_dst[t] + dst_off0 + idx * 16, v[u]);
            }
          } else {
            for (int q = 0; q < n_peers; ++q) {
              int d = p.rank + first + q;            // staggered: rank r starts with peer r+1
              if (d >= p.world) d -= p.world;
              char* base = p.dst[t][d] + dst_off0;
#pragma unroll
              for (int u = 0; u < U; ++u) {
                const int64_t idx = i + int64_t(u) * nthreads;
                if (idx < total) st_peer_v4(base + idx * 16, v[u]);
              }
            }
          }
        } else {
          int64_t off[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t idx = i + int64_t(u) * nthreads;
            off[u] = -1;
            if (idx < total) {
              int64_t row, c;
              if (p.cpr_shift >= 0) row = idx >> p.cpr_shift, c = idx & (p.cpr - 1);
              else row = idx / p.cpr, c = idx - row * p.cpr;
              v[u] = ld_once_v4(src + row * p.ld_src[t] + c * 16);
              off[u] = dst_off0 + row * p.ld_dst[t] + c * 16;
            }
          }
          if constexpr (MC) {
#pragma unroll
            for (int u = 0; u < U; ++u)
              if (off[u] >= 0) st_multicast_v4(p.mc_dst[t] + off[u], v[u]);
          } else {
            for (int q = 0; q < n_peers; ++q) {
              int d = p.rank + first + q;
              if (d >= p.world) d -= p.world;
              char* base = p.dst[t][d];
#pragma unroll
              for (int u = 0; u < U; ++u)
                if (off[u] >= 0) st_peer_v4(base + off[u], v[u]);
            }
          }
        }
      }
    }
    // slice complete on this CTA -> fence to system scope, count arrivals; the last CTA publishes
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence_system();
      const unsigned old = atomicAdd(p.counters + s, 1u);
      if (old == gridDim.x - 1) {
        p.counters[s] = 0;
        __threadfence_system();
        for (int q = 0; q < n_peers; ++q) {
          int d = p.rank + first + q;
          if (d >= p.world) d -= p.world;
          st_release_sys_u32(p.flag[d] + s, p.seq);
        }
      }
    }
  }
}

// ---- the same push with the bulk-copy (TMA) engine ---------------------------------------------------
// Contiguous shards only (ld == row_bytes on both sides).  ONE thread per CTA drives a ring of `stages`
// shared-memory tiles: cp.async.bulk global -> shared (mbarrier complete_tx) runs stages-1 tiles ahead, and
// every landed tile leaves as world-1 cp.async.bulk shared -> peer-global stores (one bulk group per
// tile; wait_group.read frees the tile).  No registers, no LSU traffic and a few dozen instructions per
// 16 KB: the CTA is 32 threads, so it fits beside the aggregation CTAs on any SM.
__device__ __forceinline__ void bulk_load(uint32_t smem, const void* g, uint32_t bytes, uint32_t bar) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem),
               "l"(g), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* g, uint32_t smem, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

struct PieceCursor {
  int s, t;
  int64_t j;     // piece index inside (slice s, tensor t); this CTA owns j = cta, cta + grid, ...
};

__global__ void __launch_bounds__(32) shard_push_tma_kernel(const __grid_constant__ PushParams p, const int chunk,
                                                            const int stages) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  if (threadIdx.x != 0) return;
  announce_start(p);
  const uint32_t tiles = tc::smem_u32(smem_raw);
  const uint32_t bars = tiles + uint32_t(stages) * uint32_t(chunk);
  for (int i = 0; i < stages; ++i) tc::mbar_init(bars + 8 * i, 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const int n_peers = p.include_self ? p.world : p.world - 1;
  const int first = p.include_self ? 0 : 1;
  const int64_t row_bytes = int64_t(p.cpr) * 16;
  auto slice_bytes = [&](int s) { return (p.slice_row[s + 1] - p.slice_row[s]) * row_bytes; };
  auto n_pieces = [&](int s) { return (slice_bytes(s) + chunk - 1) / chunk; };
  // advance a cursor to the next piece this CTA owns (or s == n_slices when none is left)
  auto settle = [&](PieceCursor& c) {
    while (c.s < p.n_slices && c.j >= n_pieces(c.s)) {
      c.j = blockIdx.x;
      if (++c.t == p.n_tensors) c.t = 0, ++c.s;
    }
  };
  PieceCursor lc{0, 0, int64_t(blockIdx.x)}, sc{0, 0, int64_t(blockIdx.x)};
  settle(lc);
  settle(sc);
  int64_t n_loaded = 0, n_stored = 0;
  auto issue_load = [&]() {
    const int64_t off = lc.j * chunk;
    const int64_t left = slice_bytes(lc.s) - off;
    const uint32_t bytes = uint32_t(left < chunk ? left : chunk);
    const int st = int(n_loaded % stages);
    bulk_load(tiles + uint32_t(st) * uint32_t(chunk), p.src[lc.t] + p.slice_row[lc.s] * row_bytes + off, bytes,
              bars + 8 * st);
    ++n_loaded;
    lc.j += gridDim.x;
    settle(lc);
  };
  while (n_loaded < stages - 1 && lc.s < p.n_slices) issue_load();

  for (int s = 0; s < p.n_slices; ++s) {
    while (sc.s == s) {
      const int64_t off = sc.j * chunk;
      const int64_t left = slice_bytes(s) - off;
      const uint32_t bytes = uint32_t(left < chunk ? left : chunk);
      const int st = int(n_stored % stages);
      tc::mbar_wait(bars + 8 * st, uint32_t(n_stored / stages) & 1u);
      const int64_t dst_off = p.slice_row[s] * row_bytes + off;
      for (int q = 0; q < n_peers; ++q) {
        int d = p.rank + first + q;
        if (d >= p.world) d -= p.world;
        bulk_store(p.dst[sc.t][d] + dst_off, tiles + uint32_t(st) * uint32_t(chunk), bytes);
      }
      bulk_commit();
      ++n_stored;
      sc.j += gridDim.x;
      settle(sc);
      bulk_wait_read_1();                       // the tile of the previous piece is free again
      if (lc.s < p.n_slices) issue_load();
    }
    bulk_wait_all();                            // every store of this slice has been performed
    asm volatile("fence.proxy.async;" ::: "memory");
    __threadfence_system();
    const unsigned old = atomicAdd(p.counters + s, 1u);
    if (old == gridDim.x - 1) {
      p.counters[s] = 0;
      __threadfence_system();
      for (int q = 0; q < n_peers; ++q) {
        int d = p.rank + first + q;
        if (d >= p.world) d -= p.world;
        st_release_sys_u32(p.flag[d] + s, p.seq);
      }
    }
  }
}

struct WaitIdx {
  int32_t v[64];
};

__global__ void wait_flags_kernel(const uint32_t* flags, const WaitIdx idx, int n, uint32_t seq, uint64_t timeout_ns,
                                  int32_t* status) {
  const int i = threadIdx.x;
  if (i >= n) return;
  const uint32_t* f = flags + idx.v[i];
  const uint64_t t0 = global_timer_ns();
  while (int32_t(ld_acquire_sys_u32(f) - seq) < 0) {
    if (global_timer_ns() - t0 > timeout_ns) {
      if (status) atomicExch(status, 1);
      break;
    }
    __nanosleep(200);
  }
}

}  // namespace pgsd

using namespace pgsd;

extern "C" size_t pgsd_sizeof_push_args(void) { return sizeof(pgsd_push_args); }

extern "C" int pgsd_shard_push(const pgsd_push_args* a, pgsd_stream_t stream) {
  PGSD_REQUIRE(a != nullptr, "shard_push: args is null");
  PGSD_REQUIRE(a->world >= 1 && a->world <= PGSD_MAX_RANKS && a->rank >= 0 && a->rank < a->world,
               "shard_push: bad world/rank %d/%d", a->world, a->rank);
  PGSD_REQUIRE(a->n_tensors == 1 || a->n_tensors == 2, "shard_push: n_tensors must be 1 or 2");
  PGSD_REQUIRE(a->n_slices >= 1 && a->n_slices <= PGSD_MAX_SLICES, "shard_push: n_slices out of range");
  PGSD_REQUIRE(a->row_bytes > 0 && a->row_bytes % 16 == 0, "shard_push: row_bytes must be a multiple of 16");
  PGSD_REQUIRE(a->counters != nullptr, "shard_push: counters is null");
  PGSD_REQUIRE(a->slice_row[0] == 0 && a->slice_row[a->n_slices] == a->n_rows, "shard_push: slices must cover the shard");
  PushParams p{};
  p.world = a->world, p.rank = a->rank, p.n_tensors = a->n_tensors, p.n_slices = a->n_slices;
  p.include_self = a->include_self ? 1 : 0;
  p.cpr = a->row_bytes / 16;
  p.cpr_shift = -1;
  for (int s = 0; s < 20; ++s)
    if ((1 << s) == p.cpr) p.cpr_shift = s;
  p.seq = a->seq;
  p.counters = a->counters;
  p.started = a->started;
  const bool mc = a->mc_dst[0] != nullptr;
  for (int t = 0; t < a->n_tensors; ++t) {
    PGSD_REQUIRE(a->src[t] != nullptr && reinterpret_cast<uintptr_t>(a->src[t]) % 16 == 0 && a->ld_src_bytes[t] % 16 == 0 &&
                     a->ld_dst_bytes[t] % 16 == 0,
                 "shard_push: rows must be 16-byte aligned");
    p.src[t] = static_cast<const char*>(a->src[t]);
    p.ld_src[t] = a->ld_src_bytes[t];
    p.ld_dst[t] = a->ld_dst_bytes[t];
    p.mc_dst[t] = static_cast<char*>(a->mc_dst[t]);
    PGSD_REQUIRE(!mc || a->mc_dst[t] != nullptr, "shard_push: mc_dst must be set for every tensor");
    for (int r = 0; r < a->world; ++r) {
      p.dst[t][r] = static_cast<char*>(a->dst[t][r]);
      PGSD_REQUIRE(mc || r == a->rank && !a->include_self || (p.dst[t][r] && reinterpret_cast<uintptr_t>(p.dst[t][r]) % 16 == 0),
                   "shard_push: dst[%d][%d] is null or unaligned", t, r);
    }
  }
  for (int s = 0; s <= a->n_slices; ++s) {
    p.slice_row[s] = a->slice_row[s];
    PGSD_REQUIRE(s == 0 || p.slice_row[s] >= p.slice_row[s - 1], "shard_push: slice_row must be non-decreasing");
  }
  for (int r = 0; r < a->world; ++r) {
    p.flag[r] = a->flag[r];
    PGSD_REQUIRE((r == a->rank && !a->include_self) || p.flag[r] != nullptr, "shard_push: flag[%d] is null", r);
  }
  int grid = a->n_ctas > 0 ? a->n_ctas : 64;
  if (grid > 2 * sm_count()) grid = 2 * sm_count();   // every CTA must be resident: the slices end on a grid-wide count
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // engine: 1 = bulk-copy (TMA) kernel, needs contiguous rows on both sides and unicast; 0 = LSU kernel
  bool contiguous = true;
  for (int t = 0; t < a->n_tensors; ++t)
    contiguous = contiguous && a->ld_src_bytes[t] == a->row_bytes && a->ld_dst_bytes[t] == a->row_bytes;
  if (a->engine == 1 && contiguous && !mc) {
    int chunk = a->chunk_bytes > 0 ? a->chunk_bytes : 16384;
    int stages = a->stages > 0 ? a->stages : 4;
    PGSD_REQUIRE(chunk % 16 == 0 && stages >= 2 && stages <= 16, "shard_push: bad chunk/stages");
    const size_t smem = size_t(chunk) * stages + 8 * stages;
    PGSD_REQUIRE(smem <= 200 * 1024, "shard_push: chunk * stages exceeds shared memory");
    static size_t smem_set = 0;
    if (smem > smem_set) {
      PGSD_CUDA(cudaFuncSetAttribute(shard_push_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
      smem_set = smem;
    }
    shard_push_tma_kernel<<<grid, 32, smem, st>>>(p, chunk, stages);
    PGSD_LAUNCH_CHECK("shard_push_tma_kernel");
    return PGSD_OK;
  }
  if (mc && contiguous) shard_push_kernel<4, true, true><<<grid, 256, 0, st>>>(p);
  else if (mc) shard_push_kernel<4, true, false><<<grid, 256, 0, st>>>(p);
  else if (contiguous) shard_push_kernel<8, false, true><<<grid, 256, 0, st>>>(p);
  else shard_push_kernel<4, false, false><<<grid, 256, 0, st>>>(p);
  PGSD_LAUNCH_CHECK("shard_push_kernel");
  return PGSD_OK;
}

// ---- content fingerprint of a device buffer (plan-cache validation, plan.py PlanCache) --------------------------
// Two wrap-around sums over the 32-bit words in ONE pass: the plain sum and a position-weighted sum (odd
// multipliers), so single-word edits, compensating edits and permutations all change the result.
__global__ void __launch_bounds__(256) fingerprint_kernel(const uint32_t* __restrict__ w, int64_t n,
                                                          unsigned long long* __restrict__ out) {
  unsigned long long s0 = 0, s1 = 0;
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  const int64_t tid = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  auto add = [&](uint32_t v, int64_t i) {
    s0 += v;
    s1 += (unsigned long long)v * (uint32_t(i) * 2654435761u | 1u);
  };
  // 128-bit loads over the 16-byte aligned body (torch allocations are), scalar loads over the tail
  const int64_t n4 = (reinterpret_cast<uintptr_t>(w) % 16 == 0) ? n / 4 : 0;
  const uint4* w4 = reinterpret_cast<const uint4*>(w);
  for (int64_t i = tid; i < n4; i += stride) {
    const uint4 v = __ldg(w4 + i);
    add(v.x, 4 * i), add(v.y, 4 * i + 1), add(v.z, 4 * i + 2), add(v.w, 4 * i + 3);
  }
  for (int64_t i = 4 * n4 + tid; i < n; i += stride) add(__ldg(w + i), i);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s0 += __shfl_xor_sync(0xffffffffu, s0, o);
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(out, s0);
    atomicAdd(out + 1, s1);
  }
}

extern "C" int pgsd_fingerprint(const void* data, int64_t n_bytes, uint64_t* out2, pgsd_stream_t stream) {
  PGSD_REQUIRE(out2 != nullptr, "fingerprint: out is null");
  const int64_t n = n_bytes / 4;
  if (n <= 0) return PGSD_OK;
  PGSD_REQUIRE(data != nullptr && reinterpret_cast<uintptr_t>(data) % 4 == 0, "fingerprint: data must be 4-byte aligned");
  int64_t grid = ceil_div<int64_t>(n, 256 * 16);
  if (grid > int64_t(sm_count()) * 8) grid = int64_t(sm_count()) * 8;
  fingerprint_kernel<<<unsigned(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const uint32_t*>(data), n, reinterpret_cast<unsigned long long*>(out2));
  PGSD_LAUNCH_CHECK("fingerprint_kernel");
  return PGSD_OK;
}

// ---- copy-engine transport (engine 2 of the Python exchange): plain peer copies + a one-word signal ----------
__global__ void signal_flags_kernel(uint32_t* flag, uint32_t seq) {
  __threadfence_system();
  st_release_sys_u32(flag, seq);
}

extern "C" int pgsd_peer_copy(void* dst, const void* src, size_t bytes, pgsd_stream_t stream) {
  if (bytes == 0) return PGSD_OK;
  PGSD_REQUIRE(dst != nullptr && src != nullptr, "peer_copy: null pointer");
  PGSD_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDefault, static_cast<cudaStream_t>(stream)));
  return PGSD_OK;
}

extern "C" int pgsd_signal_flag(uint32_t* flag, uint32_t seq, pgsd_stream_t stream) {
  PGSD_REQUIRE(flag != nullptr, "signal_flag: null pointer");
  signal_flags_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream)>>>(flag, seq);
  PGSD_LAUNCH_CHECK("signal_flags_kernel");
  return PGSD_OK;
}

extern "C" int pgsd_wait_flags(const uint32_t* flags, const int32_t* index_host, int32_t n, uint32_t seq,
                               uint64_t timeout_ns, int32_t* status, pgsd_stream_t stream) {
  PGSD_REQUIRE(flags != nullptr && (n == 0 || index_host != nullptr), "wait_flags: null pointer");
  PGSD_REQUIRE(n >= 0 && n <= 64, "wait_flags: n must be in [0, 64]");
  if (n == 0) return PGSD_OK;
  WaitIdx idx{};
  for (int i = 0; i < n; ++i) idx.v[i] = index_host[i];
  wait_flags_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(flags, idx, n, seq, timeout_ns, status);
  PGSD_LAUNCH_CHECK("wait_flags_kernel");
  return PGSD_OK;
}
