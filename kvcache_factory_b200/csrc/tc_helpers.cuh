// tcgen05 / TMA / mbarrier PTX wrappers and the softmax epilogue math shared by the window-scoring kernels
// (scores_dyn.cu).  sm_100a only.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace kvc {
namespace tcx {

constexpr int kTile = 128;                  // keys per tile, and rows per stacked query block
constexpr int kD = 128;                     // head dim covered by the tcgen05 kernels
constexpr int kHalfBytes = kTile * 64 * 2;  // one 64-column (128-byte swizzled) half of a 128 x 128 16-bit tile
constexpr int kTileBytes = 2 * kHalfBytes;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(done)
      : "r"(bar), "r"(parity)
      : "memory");
  return done != 0;
}
// Bounded spin (about a second): a wait that never completes is a protocol bug or a dead producer; trapping with a
// message beats hanging the stream forever.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try(bar, parity); ++spin) {
    if (spin > (1u << 24)) {
      printf("kvc scoring kernel: mbarrier wait timed out (block %d thread %d)\n", blockIdx.x, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tmap, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(tmap), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,"
      "%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major, 128-byte swizzle, 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t instr_desc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// ---- exp2 on the FMA pipe --------------------------------------------------------------------------------------
// 2^a for a pair of fp32 values without MUFU: Cody-Waite split a = n + f with the 1.5 * 2^23 rounding trick
// (n = round(a) lands in the low mantissa bits of t), a degree-5 minimax polynomial for 2^f on [-0.5, 0.5]
// (max relative error 1.6e-7 = 2^-22.6, the same class as ex2.approx), and the exponent put back with one integer
// shift-add per element.  Arguments below -125 are clamped (result 2^-125 instead of a flushed 0: see DESIGN.md).
// MUFU.EX2 issues at 16 lanes/clk/SM and bounds both softmax passes; routing a fraction of the exponentials through
// the half-idle FMA pipe shortens them (scripts/ubench/epi_mix.cu has the measured mixes).
__device__ __forceinline__ float2 exp2_fma2(float2 a) {
  const float kMagic = 12582912.f;
  a.x = fmaxf(a.x, -125.f);
  a.y = fmaxf(a.y, -125.f);
  const float2 t = __fadd2_rn(a, make_float2(kMagic, kMagic));
  const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __fadd2_rn(a, make_float2(-n.x, -n.y));
  float2 p = __ffma2_rn(make_float2(0.00132764713f, 0.00132764713f), f, make_float2(0.00967554096f, 0.00967554096f));
  p = __ffma2_rn(p, f, make_float2(0.0555071309f, 0.0555071309f));
  p = __ffma2_rn(p, f, make_float2(0.240221202f, 0.240221202f));
  p = __ffma2_rn(p, f, make_float2(0.693146944f, 0.693146944f));
  p = __ffma2_rn(p, f, make_float2(1.00000012f, 1.00000012f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

// Lazy running maximum of pass 1: the exponentials of a row are taken relative to a reference m_ref that is only
// raised when a 32-key chunk exceeds it by more than kTau (a warp-uniform, rare branch); (m_ref, Z) still
// represents Z * e^m_ref exactly.
constexpr float kTau = 6.0f;

// Pass 1, one chunk of 32 keys of this thread's query row: the reference's two roundings (GEMM result -> dtype,
// * 1/sqrt(d) in fp32 -> dtype), lazy max, Z += sum exp(x - m_ref).  POLY of the 16 element pairs take their
// exponentials on the FMA pipe.
template <typename T, bool MASK, int POLY>
__device__ __forceinline__ void p1_chunk(const uint32_t (&v)[32], float scale, float& m_ref, float2 (&zacc)[4], int key0,
                                         int last_key) {
  const float2 scale2 = make_float2(scale, scale);
  const float2 l2e2 = make_float2(kLog2e, kLog2e);
  uint32_t r2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 f = __fmul2_rn(make_float2(T::rnd(__uint_as_float(v[2 * i])), T::rnd(__uint_as_float(v[2 * i + 1]))), scale2);
    r2[i] = T::pack(f.x, f.y);
  }
  if (MASK) {  // chunk touches the w x w causal block or the ragged end
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int key = key0 + 2 * i;
      if (key > last_key) r2[i] = (r2[i] & 0xffff0000u) | (T::kNegInfPair & 0xffffu);
      if (key + 1 > last_key) r2[i] = (r2[i] & 0x0000ffffu) | (T::kNegInfPair & 0xffff0000u);
    }
  }
  uint32_t pm = r2[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) pm = T::max2(pm, r2[i]);
  const float cm = fmaxf(T::lo(pm), T::hi(pm));
  if (__any_sync(0xffffffffu, cm > m_ref + kTau)) {
    const float m_new = cm > m_ref + kTau ? cm : m_ref;
    const float sc = fast_exp2((m_ref - m_new) * kLog2e);
    const float2 sc2 = make_float2(sc, sc);
#pragma unroll
    for (int a = 0; a < 4; ++a) zacc[a] = __fmul2_rn(zacc[a], sc2);
    m_ref = m_new;
  }
  const float2 ml2 = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 a = __ffma2_rn(make_float2(T::lo(r2[i]), T::hi(r2[i])), l2e2, ml2);
    const bool poly = POLY > 0 && ((i * POLY) / 16 != ((i + 1) * POLY) / 16);
    const float2 e = poly ? exp2_fma2(a) : make_float2(fast_exp2(a.x), fast_exp2(a.y));
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], e);
  }
}

// Pass 2, one chunk of 32 stacked query rows of this thread's key: the same two roundings,
// p = exp2(x*log2e + c_row) rounded to dtype (softmax output), summed per head in fp32.
template <typename T, int W, int POLY>
__device__ __forceinline__ void p2_chunk(const uint32_t (&v)[32], float scale, const float* cr, float2 (&acc)[32 / W]) {
  const float2 scale2 = make_float2(scale, scale);
  const float2 l2e2 = make_float2(kLog2e, kLog2e);
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 c4 = *reinterpret_cast<const float4*>(cr + q4 * 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * q4 + h;
      float2 f = __fmul2_rn(make_float2(T::rnd(__uint_as_float(v[2 * i])), T::rnd(__uint_as_float(v[2 * i + 1]))), scale2);
      f = __ffma2_rn(make_float2(T::rnd(f.x), T::rnd(f.y)), l2e2, h == 0 ? make_float2(c4.x, c4.y) : make_float2(c4.z, c4.w));
      const bool poly = POLY > 0 && ((i * POLY) / 16 != ((i + 1) * POLY) / 16);
      const float2 e0 = poly ? exp2_fma2(f) : make_float2(fast_exp2(f.x), fast_exp2(f.y));
      const float2 e = make_float2(T::rnd(e0.x), T::rnd(e0.y));
      acc[(2 * i) / W] = __fadd2_rn(acc[(2 * i) / W], e);
    }
  }
}

// ---- 16-column halves (scores_dyn.cu): the tcgen05.ld of one half is in flight while the other half is computed ----
// Pass 1, 16 keys of this thread's query row; same math as p1_chunk (the lazy reference is checked per half).
template <typename T, bool MASK, int POLY>
__device__ __forceinline__ void p1_half(const uint32_t (&v)[16], float scale, float& m_ref, float2 (&zacc)[4], int key0,
                                        int last_key) {
  const float2 scale2 = make_float2(scale, scale);
  const float2 l2e2 = make_float2(kLog2e, kLog2e);
  uint32_t r2[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 f = __fmul2_rn(make_float2(T::rnd(__uint_as_float(v[2 * i])), T::rnd(__uint_as_float(v[2 * i + 1]))), scale2);
    r2[i] = T::pack(f.x, f.y);
  }
  if (MASK) {  // half touches the w x w causal block or the ragged end
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = key0 + 2 * i;
      if (key > last_key) r2[i] = (r2[i] & 0xffff0000u) | (T::kNegInfPair & 0xffffu);
      if (key + 1 > last_key) r2[i] = (r2[i] & 0x0000ffffu) | (T::kNegInfPair & 0xffff0000u);
    }
  }
  uint32_t pm = T::max2(T::max2(T::max2(r2[0], r2[1]), T::max2(r2[2], r2[3])), T::max2(T::max2(r2[4], r2[5]), T::max2(r2[6], r2[7])));
  const float cm = fmaxf(T::lo(pm), T::hi(pm));
  if (__any_sync(0xffffffffu, cm > m_ref + kTau)) {
    const float m_new = cm > m_ref + kTau ? cm : m_ref;
    const float sc = fast_exp2((m_ref - m_new) * kLog2e);
    const float2 sc2 = make_float2(sc, sc);
#pragma unroll
    for (int a = 0; a < 4; ++a) zacc[a] = __fmul2_rn(zacc[a], sc2);
    m_ref = m_new;
  }
  const float2 ml2 = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float2 a = __ffma2_rn(make_float2(T::lo(r2[i]), T::hi(r2[i])), l2e2, ml2);
    const bool poly = POLY > 0 && ((i * POLY) / 8 != ((i + 1) * POLY) / 8);
    const float2 e = poly ? exp2_fma2(a) : make_float2(fast_exp2(a.x), fast_exp2(a.y));
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], e);
  }
}

// Pass 2, 16 stacked query rows (columns col0 .. col0+15 of the warp's 32) of this thread's key; acc[] spans the 32.
template <typename T, int W, int POLY>
__device__ __forceinline__ void p2_half(const uint32_t (&v)[16], float scale, const float* cr, int col0,
                                        float2 (&acc)[32 / W]) {
  const float2 scale2 = make_float2(scale, scale);
  const float2 l2e2 = make_float2(kLog2e, kLog2e);
#pragma unroll
  for (int q4 = 0; q4 < 4; ++q4) {
    const float4 c4 = *reinterpret_cast<const float4*>(cr + q4 * 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * q4 + h;
      float2 f = __fmul2_rn(make_float2(T::rnd(__uint_as_float(v[2 * i])), T::rnd(__uint_as_float(v[2 * i + 1]))), scale2);
      f = __ffma2_rn(make_float2(T::rnd(f.x), T::rnd(f.y)), l2e2, h == 0 ? make_float2(c4.x, c4.y) : make_float2(c4.z, c4.w));
      const bool poly = POLY > 0 && ((i * POLY) / 8 != ((i + 1) * POLY) / 8);
      const float2 e0 = poly ? exp2_fma2(f) : make_float2(fast_exp2(f.x), fast_exp2(f.y));
      const float2 e = make_float2(T::rnd(e0.x), T::rnd(e0.y));
      acc[(col0 + 2 * i) / W] = __fadd2_rn(acc[(col0 + 2 * i) / W], e);
    }
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// K [B,Hkv,L,d] with element strides -> 4-D tensor map, box = 64 columns x 128 keys, 128-byte swizzle
inline int make_k_tensor_map(CUtensorMap* tmap, const void* k, int64_t k_sb, int64_t k_sh, int64_t k_sl, int B, int Hkv,
                             int L, int d, int dtype) {
  EncodeTiledFn enc = encode_fn();
  KVC_REQUIRE(enc != nullptr, KVC_ERR_DEVICE, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)d, (cuuint64_t)L, (cuuint64_t)Hkv, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)k_sl * 2, (cuuint64_t)k_sh * 2, (cuuint64_t)k_sb * 2};
  if (Hkv == 1) gstr[1] = gstr[0] * (cuuint64_t)L;
  if (B == 1) gstr[2] = gstr[0] * (cuuint64_t)L * (cuuint64_t)Hkv;
  const cuuint32_t box[4] = {64, (cuuint32_t)kTile, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(tmap, dtype == KVC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                          4, const_cast<void*>(k), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVC_REQUIRE(cr == CUDA_SUCCESS, KVC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
  return 0;
}

}  // namespace tcx
}  // namespace kvc
