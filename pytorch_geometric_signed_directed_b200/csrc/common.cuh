// Shared helpers for the pgsd_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/pgsd_b200.h"

namespace pgsd {

// ---- per-thread error message -------------------------------------------------------------
inline char* err_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(err_buf(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define PGSD_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return pgsd::fail(PGSD_ERR_CUDA, "%s failed at %s:%d: %s", #call, __FILE__,        \
                        __LINE__, cudaGetErrorString(e__));                               \
  } while (0)

#define PGSD_LAUNCH_CHECK(name)                                                           \
  do {                                                                                    \
    cudaError_t e__ = cudaGetLastError();                                                 \
    if (e__ != cudaSuccess)                                                               \
      return pgsd::fail(PGSD_ERR_CUDA, "launch of %s failed: %s", name,                  \
                        cudaGetErrorString(e__));                                         \
  } while (0)

#define PGSD_REQUIRE(cond, ...)                                                           \
  do {                                                                                    \
    if (!(cond)) return pgsd::fail(PGSD_ERR_INVALID, __VA_ARGS__);                        \
  } while (0)

inline int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <typename T>
inline T ceil_div(T a, T b) {
  return (a + b - 1) / b;
}
inline size_t align_up(size_t v, size_t a = 256) { return (v + a - 1) / a * a; }

// ---- L2 cache policies (createpolicy) + hinted loads / stores ------------------------------
// Gathered feature rows are re-referenced ~deg times across the kernel and should survive in
// the 126 MB L2; index/value streams and outputs are touched once and should not displace
// them.
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__device__ __forceinline__ int ld_stream_i32(const int* p, uint64_t pol) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.b32 %0, [%1], %2;"
               : "=r"(r)
               : "l"(p), "l"(pol));
  return r;
}
__device__ __forceinline__ float ld_stream_f32(const float* p, uint64_t pol) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;"
               : "=f"(r)
               : "l"(p), "l"(pol));
  return r;
}

// 128-bit gather with an L2 policy.
__device__ __forceinline__ float4 ld_gather_v4(const void* p, uint64_t pol) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}
// 256-bit gather (sm_100 LDG.256) with L2 evict_last.
struct float8 {
  float v[8];
};
__device__ __forceinline__ float8 ld_gather_v8(const void* p) {
  float8 r;
  asm volatile(
      "ld.global.nc.L1::no_allocate.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(r.v[0]), "=f"(r.v[1]), "=f"(r.v[2]), "=f"(r.v[3]), "=f"(r.v[4]), "=f"(r.v[5]),
        "=f"(r.v[6]), "=f"(r.v[7])
      : "l"(p));
  return r;
}
// Gathers into caller-named registers (raw 32-bit words; bf16 pairs are expanded at FMA time so that no
// ALU work that waits on a load sits between the U * n_ops gathers of a batch).  Lanes without an entry
// skip the load (`if (ok) load; else zero;` -- ptxas if-converts it to a predicated LDG over zeroed
// registers).  Do NOT replace the predication by a dummy address: every idle lane of the GPU then hits
// the same L2 sector and the kernels run 2x slower (profiles/README.md, session 14).
__device__ __forceinline__ void ld_gather_v4_to(const void* p, uint64_t pol, float (&w)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;"
               : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3])
               : "l"(p), "l"(pol));
}
// no L2 policy operand (the immediate .L2::evict_* qualifiers are accepted only on 256-bit loads;
// the policy of the gathers measured as a no-op, profiles/r01_sweep_l2policy.jsonl)
__device__ __forceinline__ void ld_gather_v4_plain_to(const void* p, float (&w)[4]) {
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3])
               : "l"(p));
}
__device__ __forceinline__ void ld_gather_v8_to(const void* p, float (&w)[8]) {
  asm volatile(
      "ld.global.nc.L1::no_allocate.L2::evict_last.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
      : "=f"(w[0]), "=f"(w[1]), "=f"(w[2]), "=f"(w[3]), "=f"(w[4]), "=f"(w[5]), "=f"(w[6]), "=f"(w[7])
      : "l"(p));
}
__device__ __forceinline__ int ld_once_i32(const int* p) {
  int r;
  asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_once_f32(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
// base + index * stride with ONE IMAD.WIDE.U32 (row strides are < 4 GiB)
__device__ __forceinline__ const char* row_addr(const char* base, int index, uint32_t stride_bytes) {
  return base + uint64_t(uint32_t(index)) * stride_bytes;
}
__device__ __forceinline__ void st_stream_v4(void* p, float4 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v4.f32 [%0], {%1,%2,%3,%4}, %5;" ::"l"(p),
               "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol)
               : "memory");
}
__device__ __forceinline__ void st_stream_v2(void* p, float2 v, uint64_t pol) {
  asm volatile("st.global.L1::no_allocate.L2::cache_hint.v2.f32 [%0], {%1,%2}, %3;" ::"l"(p),
               "f"(v.x), "f"(v.y), "l"(pol)
               : "memory");
}

// bf16 <-> fp32 packing
__device__ __forceinline__ float bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

}  // namespace pgsd
