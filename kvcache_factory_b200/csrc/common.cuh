// Shared device/host helpers for libkvc_b200 (sm_100a only).
#pragma once

#include <atomic>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdarg>
#include <cstdio>

#include "../../include/kvc_b200.h"

namespace kvc {

// ---- error plumbing (thread-local message behind kvc_last_error) ---------------------------------
char* last_error_buffer();
int fail(int code, const char* fmt, ...);

#define KVC_CUDA_TRY(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess)                                                                  \
      return ::kvc::fail((int)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__),  \
                         __FILE__, __LINE__);                                                \
  } while (0)

#define KVC_REQUIRE(cond, code, ...) \
  do {                               \
    if (!(cond)) return ::kvc::fail((code), __VA_ARGS__); \
  } while (0)

void count_launch();  // bumps the counter behind kvc_launch_count()

inline int launch_status(const char* what) {
  count_launch();
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail((int)e, "launch of %s failed: %s", what, cudaGetErrorString(e));
  return 0;
}

int sm_count();  // cached per device

// ---- model dtype traits: rounding points of the reference live here ----------------------------
// Packed pair layout: low 16 bits = first (even column) element, high 16 bits = second.
struct Bf16 {
  static constexpr int kDtype = KVC_BF16;
  __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static __forceinline__ float lo(uint32_t p) { return __uint_as_float(p << 16); }
  __device__ static __forceinline__ float hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
  // round to bf16, result kept in fp32 format: one F2FP with a zero partner (no unpack needed)
  __device__ static __forceinline__ float rnd(float x) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(0.f));
    return __uint_as_float(r);
  }
  __device__ static __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __device__ static __forceinline__ uint16_t round1(float x) {
    __nv_bfloat16 v = __float2bfloat16_rn(x);
    return *reinterpret_cast<uint16_t*>(&v);
  }
  __device__ static __forceinline__ float to_float(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
  static constexpr uint32_t kNegInfPair = 0xff80ff80u;
};

struct Fp16 {
  static constexpr int kDtype = KVC_FP16;
  __device__ static __forceinline__ uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static __forceinline__ float lo(uint32_t p) {
    return __low2float(*reinterpret_cast<__half2*>(&p));
  }
  __device__ static __forceinline__ float hi(uint32_t p) {
    return __high2float(*reinterpret_cast<__half2*>(&p));
  }
  __device__ static __forceinline__ float rnd(float x) { return __half2float(__float2half_rn(x)); }
  __device__ static __forceinline__ uint32_t max2(uint32_t a, uint32_t b) {
    __half2 r = __hmax2(*reinterpret_cast<__half2*>(&a), *reinterpret_cast<__half2*>(&b));
    return *reinterpret_cast<uint32_t*>(&r);
  }
  __device__ static __forceinline__ uint16_t round1(float x) {
    __half v = __float2half_rn(x);
    return *reinterpret_cast<uint16_t*>(&v);
  }
  __device__ static __forceinline__ float to_float(uint16_t b) {
    return __half2float(*reinterpret_cast<__half*>(&b));
  }
  static constexpr uint32_t kNegInfPair = 0xfc00fc00u;
};

// Order-preserving map of a 16-bit float pattern (either format) onto uint16: larger value ->
// larger key; -0 < +0, NaNs with the sign bit clear sort above +inf (torch.topk treats NaN as
// the largest value).
__device__ __forceinline__ uint32_t float16_bits_to_key(uint32_t bits) {
  return (bits & 0x8000u) ? (~bits & 0xffffu) : (bits | 0x8000u);
}
__device__ __forceinline__ uint32_t key_to_float16_bits(uint32_t key) {
  return (key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu);
}

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 16-byte global accesses that do not pollute L1 (streaming rows are touched once).
__device__ __forceinline__ int4 ld_stream16(const void* p) {
  int4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void st_stream16(void* p, const int4& v) {
  asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
               "r"(v.z), "r"(v.w)
               : "memory");
}

constexpr float kLog2e = 1.4426950408889634f;

// ---- developer timeline (`make TRACE=1`): per launch, the earliest CTA start and the latest CTA end in
// globaltimer ns; kvc_debug_timeline() reads them back.  tl == nullptr in the shipped library.
#ifdef KVC_TRACE
unsigned long long* timeline_slot(int kind);  // api.cu: next entry {start, ~end}, both initialised to all-ones
__device__ __forceinline__ void tl_start(unsigned long long* tl) {
  if (tl && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(tl, t); }
}
__device__ __forceinline__ void tl_end(unsigned long long* tl) {
  if (tl && threadIdx.x == 0) { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); atomicMin(tl + 1, ~t); }
}
#define KVC_TL_SLOT(kind) timeline_slot(kind)
#else
__device__ __forceinline__ void tl_start(unsigned long long*) {}
__device__ __forceinline__ void tl_end(unsigned long long*) {}
#define KVC_TL_SLOT(kind) nullptr
#endif

// s / d, correctly rounded, from a precomputed reciprocal: quotient estimate, exact remainder, one correction
// (Markstein).  Tiny numerators are scaled by 2^64 first so that the remainder cannot underflow (the scaling is
// exact; only a subnormal quotient is rounded a second time).  No `/` anywhere: a zero numerator -- underflowed
// probabilities, the padded tail of a row -- sends the compiler's division down its special-case subroutine, and the
// one warp holding the tail then held the whole cluster at its first barrier for 2 us (scripts/trace_sel.py,
// n = 32736 against 32768: 25.3k against 21.3k cycles per row).
__device__ __forceinline__ float div_exact(float s, float d, float rcp) {
  const bool tiny = fabsf(s) < 1e-30f;
  const float up = tiny ? 0x1p64f : 1.0f, down = tiny ? 0x1p-64f : 1.0f;
  const float a = s * up;
  const float q = a * rcp;
  const float r = fmaf(-q, d, a);
  return fmaf(r, rcp, q) * down;
}

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per (call site = kernel instance, device) and only when the
// request grows: the call costs ~1 us of host time, which the eager drop-in path of short prompts is bound by.
// The call-site statics live in function templates, i.e. one set per kernel instantiation.
#define KVC_SET_MAX_SMEM(kern, bytes)                                                                          \
  do {                                                                                                         \
    static std::atomic<int> kvc_smem_set_[64];                                                                 \
    int kvc_dev_ = 0;                                                                                          \
    if (cudaGetDevice(&kvc_dev_) != cudaSuccess || kvc_dev_ < 0 || kvc_dev_ >= 64) kvc_dev_ = 0;               \
    if (kvc_smem_set_[kvc_dev_].load(std::memory_order_relaxed) < (int)(bytes)) {                              \
      KVC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));     \
      kvc_smem_set_[kvc_dev_].store((int)(bytes), std::memory_order_relaxed);                                  \
    }                                                                                                          \
  } while (0)

}  // namespace kvc
