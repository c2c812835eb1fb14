// Compute-only cost of candidate scoring-epilogue instruction mixes (run on B200): cycles per 32-element chunk per
// SM sub-partition with 4 epilogue warps per scheduler, values already in registers (no TMEM / barriers).
//   p1_cur   : pass 1 as shipped in round 1 (pack, packed max, lazy reference, unpack, FFMA2, 2 MUFU, FADD2)
//   p1_nomax : no packed max; reference fixed, overflow checked on the accumulated sum once per chunk
//   p1_poly<F>: p1_nomax with F of 16 element pairs through the FMA-pipe polynomial instead of MUFU
//   p2_cur   : pass 2 as shipped;  p2_poly<F>: F of 16 pairs through the polynomial
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bin/epi_mix epi_mix.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#define ITER 1024
constexpr float kLog2e = 1.4426950408889634f;
__device__ __forceinline__ float ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rnd(float x) { uint32_t r; asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x), "f"(0.f)); return __uint_as_float(r); }
__device__ __forceinline__ uint32_t pack(float lo, float hi) { __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi); return *reinterpret_cast<uint32_t*>(&v); }
__device__ __forceinline__ float lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float hi(uint32_t p) { return __uint_as_float(p & 0xffff0000u); }
__device__ __forceinline__ uint32_t max2(uint32_t a, uint32_t b) { __nv_bfloat162 r = __hmax2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b)); return *reinterpret_cast<uint32_t*>(&r); }

// 2^a for a pair, on the FMA pipe: Cody-Waite split with the 1.5*2^23 trick + degree-DEG minimax polynomial on
// [-0.5, 0.5], exponent put back with one integer shift-add per element.
template <int DEG>
__device__ __forceinline__ float2 exp2_poly2(float2 a) {
  const float kMagic = 12582912.f;
  a.x = fmaxf(a.x, -125.f); a.y = fmaxf(a.y, -125.f);
  const float2 t = __fadd2_rn(a, make_float2(kMagic, kMagic));
  const float2 n = __fadd2_rn(t, make_float2(-kMagic, -kMagic));
  const float2 f = __fadd2_rn(a, make_float2(-n.x, -n.y));
  float2 p;
  if (DEG == 3) {
    p = __ffma2_rn(make_float2(0.0551716611f, 0.0551716611f), f, make_float2(0.242611125f, 0.242611125f));
    p = __ffma2_rn(p, f, make_float2(0.693260968f, 0.693260968f));
    p = __ffma2_rn(p, f, make_float2(0.999928057f, 0.999928057f));
  } else if (DEG == 4) {
    p = __ffma2_rn(make_float2(0.00957010128f, 0.00957010128f), f, make_float2(0.0559178591f, 0.0559178591f));
    p = __ffma2_rn(p, f, make_float2(0.240247443f, 0.240247443f));
    p = __ffma2_rn(p, f, make_float2(0.693121791f, 0.693121791f));
    p = __ffma2_rn(p, f, make_float2(0.999999285f, 0.999999285f));
  } else {
    p = __ffma2_rn(make_float2(0.00132764713f, 0.00132764713f), f, make_float2(0.00967554096f, 0.00967554096f));
    p = __ffma2_rn(p, f, make_float2(0.0555071309f, 0.0555071309f));
    p = __ffma2_rn(p, f, make_float2(0.240221202f, 0.240221202f));
    p = __ffma2_rn(p, f, make_float2(0.693146944f, 0.693146944f));
    p = __ffma2_rn(p, f, make_float2(1.00000012f, 1.00000012f));
  }
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23)));
}

__device__ __forceinline__ void p1_cur(const uint32_t (&v)[32], float scale, float& m_ref, float2 (&zacc)[4]) {
  const float2 scale2 = make_float2(scale, scale), l2e2 = make_float2(kLog2e, kLog2e);
  uint32_t r2[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 f = __fmul2_rn(make_float2(rnd(__uint_as_float(v[2 * i])), rnd(__uint_as_float(v[2 * i + 1]))), scale2);
    r2[i] = pack(f.x, f.y);
  }
  uint32_t pm = r2[0];
#pragma unroll
  for (int i = 1; i < 16; ++i) pm = max2(pm, r2[i]);
  const float cm = fmaxf(lo(pm), hi(pm));
  if (__any_sync(0xffffffffu, cm > m_ref + 6.f)) {
    const float m_new = cm > m_ref + 6.f ? cm : m_ref;
    const float sc = ex2((m_ref - m_new) * kLog2e);
#pragma unroll
    for (int a = 0; a < 4; ++a) zacc[a] = __fmul2_rn(zacc[a], make_float2(sc, sc));
    m_ref = m_new;
  }
  const float2 ml2 = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 a = __ffma2_rn(make_float2(lo(r2[i]), hi(r2[i])), l2e2, ml2);
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], make_float2(ex2(a.x), ex2(a.y)));
  }
}

// round-1 math, but the exponentials are taken SPECULATIVELY with the current reference before the packed-max vote
// resolves (the vote then only guards a rare redo): the max tree leaves the critical path, results are bit-identical
template <int F, int DEG>
__device__ __forceinline__ void p1_spec(const uint32_t (&v)[32], float scale, float& m_ref, float2 (&zacc)[4]) {
  const float2 scale2 = make_float2(scale, scale), l2e2 = make_float2(kLog2e, kLog2e);
  uint32_t r2[16];
  float2 zs[4] = {zacc[0], zacc[1], zacc[2], zacc[3]};
  const float2 ml2 = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
  uint32_t pm = 0xff80ff80u;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float2 f = __fmul2_rn(make_float2(rnd(__uint_as_float(v[2 * i])), rnd(__uint_as_float(v[2 * i + 1]))), scale2);
    r2[i] = pack(f.x, f.y);
    pm = max2(pm, r2[i]);
    const float2 a = __ffma2_rn(make_float2(lo(r2[i]), hi(r2[i])), l2e2, ml2);
    const bool poly = F > 0 && ((i * F) / 16 != ((i + 1) * F) / 16);
    const float2 e = poly ? exp2_poly2<DEG>(a) : make_float2(ex2(a.x), ex2(a.y));
    zs[i & 3] = __fadd2_rn(zs[i & 3], e);
  }
  const float cm = fmaxf(lo(pm), hi(pm));
  if (__any_sync(0xffffffffu, cm > m_ref + 6.f)) {   // rare: redo the chunk with the raised reference
    const float m_new = cm > m_ref + 6.f ? cm : m_ref;
    const float sc = ex2((m_ref - m_new) * kLog2e);
#pragma unroll
    for (int a = 0; a < 4; ++a) zs[a] = __fmul2_rn(zacc[a], make_float2(sc, sc));
    m_ref = m_new;
    const float2 ml = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float2 a = __ffma2_rn(make_float2(lo(r2[i]), hi(r2[i])), l2e2, ml);
      zs[i & 3] = __fadd2_rn(zs[i & 3], make_float2(ex2(a.x), ex2(a.y)));
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) zacc[a] = zs[a];
}

// F pairs of 16 through the polynomial (spread evenly), the rest through MUFU
template <int F, int DEG>
__device__ __forceinline__ void p1_new(const uint32_t (&v)[32], float scale, float& m_ref, float2 (&zacc)[4]) {
  const float2 scale2 = make_float2(scale, scale), l2e2 = make_float2(kLog2e, kLog2e);
  const float2 ml2 = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    float2 f = __fmul2_rn(make_float2(rnd(__uint_as_float(v[2 * i])), rnd(__uint_as_float(v[2 * i + 1]))), scale2);
    f = __ffma2_rn(make_float2(rnd(f.x), rnd(f.y)), l2e2, ml2);
    const bool poly = F > 0 && ((i * F) / 16 != ((i + 1) * F) / 16);
    const float2 e = poly ? exp2_poly2<DEG>(f) : make_float2(ex2(f.x), ex2(f.y));
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], e);
  }
  // overflow guard: once per chunk on the accumulated sums (terms are positive, sums only grow)
  const float zs = (zacc[0].x + zacc[1].x) + (zacc[2].x + zacc[3].x) + (zacc[0].y + zacc[1].y) + (zacc[2].y + zacc[3].y);
  if (__any_sync(0xffffffffu, zs > 1e30f)) {
    const float sh = zs > 1e30f ? 64.f : 0.f;
    m_ref += sh * 0.6931471805599453f;
    const float sc = ex2(-sh);
#pragma unroll
    for (int a = 0; a < 4; ++a) zacc[a] = __fmul2_rn(zacc[a], make_float2(sc, sc));
  }
}

template <int F, int DEG>
__device__ __forceinline__ void p2_new(const uint32_t (&v)[32], float scale, const float* cr, float2& acc) {
  const float2 scale2 = make_float2(scale, scale), l2e2 = make_float2(kLog2e, kLog2e);
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 c4 = *reinterpret_cast<const float4*>(cr + q4 * 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * q4 + h;
      float2 f = __fmul2_rn(make_float2(rnd(__uint_as_float(v[2 * i])), rnd(__uint_as_float(v[2 * i + 1]))), scale2);
      f = __ffma2_rn(make_float2(rnd(f.x), rnd(f.y)), l2e2, h == 0 ? make_float2(c4.x, c4.y) : make_float2(c4.z, c4.w));
      const bool poly = F > 0 && ((i * F) / 16 != ((i + 1) * F) / 16);
      const float2 e0 = poly ? exp2_poly2<DEG>(f) : make_float2(ex2(f.x), ex2(f.y));
      acc = __fadd2_rn(acc, make_float2(rnd(e0.x), rnd(e0.y)));
    }
  }
}

// packed variants: the model-dtype roundings as ONE F2FP per pair (bf16x2) + unpack split over the FMA pipe
// (IMAD.U32 x,0x10000) and the ALU pipe (LOP3), instead of one F2FP per element
template <int F, int DEG>
__device__ __forceinline__ void p1_packed(const uint32_t (&v)[32], float scale, float& m_ref, float2 (&zacc)[4]) {
  const float2 scale2 = make_float2(scale, scale), l2e2 = make_float2(kLog2e, kLog2e);
  const float2 ml2 = make_float2(-m_ref * kLog2e, -m_ref * kLog2e);
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const uint32_t a = pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
    float2 f = __fmul2_rn(make_float2(lo(a), hi(a)), scale2);
    const uint32_t b = pack(f.x, f.y);
    f = __ffma2_rn(make_float2(lo(b), hi(b)), l2e2, ml2);
    const bool poly = F > 0 && ((i * F) / 16 != ((i + 1) * F) / 16);
    const float2 e = poly ? exp2_poly2<DEG>(f) : make_float2(ex2(f.x), ex2(f.y));
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], e);
  }
  const float zs = (zacc[0].x + zacc[1].x) + (zacc[2].x + zacc[3].x) + (zacc[0].y + zacc[1].y) + (zacc[2].y + zacc[3].y);
  if (__any_sync(0xffffffffu, zs > 1e30f)) {
    const float sh = zs > 1e30f ? 64.f : 0.f;
    m_ref += sh * 0.6931471805599453f;
    const float sc = ex2(-sh);
#pragma unroll
    for (int a = 0; a < 4; ++a) zacc[a] = __fmul2_rn(zacc[a], make_float2(sc, sc));
  }
}
template <int F, int DEG>
__device__ __forceinline__ void p2_packed(const uint32_t (&v)[32], float scale, const float* cr, float2& acc) {
  const float2 scale2 = make_float2(scale, scale), l2e2 = make_float2(kLog2e, kLog2e);
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 c4 = *reinterpret_cast<const float4*>(cr + q4 * 4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * q4 + h;
      const uint32_t a = pack(__uint_as_float(v[2 * i]), __uint_as_float(v[2 * i + 1]));
      float2 f = __fmul2_rn(make_float2(lo(a), hi(a)), scale2);
      const uint32_t b = pack(f.x, f.y);
      f = __ffma2_rn(make_float2(lo(b), hi(b)), l2e2, h == 0 ? make_float2(c4.x, c4.y) : make_float2(c4.z, c4.w));
      const bool poly = F > 0 && ((i * F) / 16 != ((i + 1) * F) / 16);
      const float2 e0 = poly ? exp2_poly2<DEG>(f) : make_float2(ex2(f.x), ex2(f.y));
      const uint32_t c = pack(e0.x, e0.y);
      acc = __fadd2_rn(acc, make_float2(lo(c), hi(c)));
    }
  }
}

template <int MODE, int F, int DEG>
__global__ void __launch_bounds__(512, 1) k(float* out, const uint32_t* in, int nthr) {
  __shared__ float cr[64];
  extern __shared__ uint4 vsm[];   // [8][512] uint4: thread t reads vsm[j*512 + t], conflict-free 16-byte accesses
  if (threadIdx.x < 64) cr[threadIdx.x] = -3.f - 0.01f * threadIdx.x;
  for (int j = 0; j < 8; ++j) {
    const uint32_t* p = in + ((threadIdx.x * 32 + j * 4) & 4095);
    vsm[j * 512 + threadIdx.x] = make_uint4(p[0], p[1], p[2], p[3]);
  (void)nthr;
  }
  __syncthreads();
  float m_ref = 2.f;
  float2 zacc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
  float2 acc = make_float2(0.f, 0.f);
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
    uint32_t v[32];
#pragma unroll
    for (int j = 0; j < 8; ++j) {   // stands in for the tcgen05.ld of the real kernel: the values are new every iteration
      const uint32_t sa = (uint32_t)__cvta_generic_to_shared(&vsm[j * 512 + threadIdx.x]);
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3]) : "r"(sa));
    }
    const float* crp = cr + (it & 1) * 32;
    if (MODE == 0) p1_cur(v, 0.0883883461f, m_ref, zacc);
    if (MODE == 1) p1_new<F, DEG>(v, 0.0883883461f, m_ref, zacc);
    if (MODE == 2) p2_new<F, DEG>(v, 0.0883883461f, crp, acc);
    if (MODE == 3) p1_packed<F, DEG>(v, 0.0883883461f, m_ref, zacc);
    if (MODE == 4) p2_packed<F, DEG>(v, 0.0883883461f, crp, acc);
    if (MODE == 5) p1_spec<F, DEG>(v, 0.0883883461f, m_ref, zacc);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = zacc[0].x + zacc[1].y + zacc[2].x + zacc[3].y + acc.x + acc.y + m_ref;
}

template <int MODE, int F, int DEG>
void run(const char* name) {
  float* d; uint32_t* in; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&in, 4096 * 4);
  uint32_t h[4096];
  for (int i = 0; i < 4096; ++i) { float x = ((i * 2654435761u) >> 8) % 2000 * 0.02f - 20.f; h[i] = *reinterpret_cast<uint32_t*>(&x); }
  cudaMemcpy(in, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(k<MODE, F, DEG>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  double res[2];
  for (int w = 0; w < 2; ++w) {
    const int thr = w == 0 ? 128 : 512;   // 1 or 4 warps per SM sub-partition
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE, F, DEG><<<148, thr, 65536>>>(d, in, thr); cudaDeviceSynchronize();
    cudaEventRecord(e0); k<MODE, F, DEG><<<148, thr, 65536>>>(d, in, thr); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    res[w] = ms * 1e-3 * 1.965e9 / ITER;   // cycles per 32-element chunk of ONE warp
  }
  printf("%-36s 1 warp/SMSP: %5.0f clk per chunk (latency of a lone warp) | 4 warps/SMSP: %5.0f clk per chunk each = %5.0f clk per 128x128 tile\n",
         name, res[0], res[1], res[1]);
  cudaFree(d); cudaFree(in);
}

int main() {
  run<0, 0, 4>("p1 round-1 (packed max, lazy ref)");
  run<5, 0, 4>("p1 speculative exps, 0/16 poly");
  run<5, 4, 5>("p1 speculative exps, 4/16 deg5");
  run<5, 6, 5>("p1 speculative exps, 6/16 deg5");
  run<1, 0, 4>("p1 nomax, 0/16 poly");
  run<1, 4, 4>("p1 nomax, 4/16 poly deg4");
  run<1, 5, 4>("p1 nomax, 5/16 poly deg4");
  run<1, 6, 4>("p1 nomax, 6/16 poly deg4");
  run<1, 7, 4>("p1 nomax, 7/16 poly deg4");
  run<1, 6, 3>("p1 nomax, 6/16 poly deg3");
  run<1, 6, 5>("p1 nomax, 6/16 poly deg5");
  run<3, 0, 4>("p1 packed roundings, 0/16 poly");
  run<3, 4, 4>("p1 packed roundings, 4/16 deg4");
  run<3, 6, 4>("p1 packed roundings, 6/16 deg4");
  run<3, 8, 4>("p1 packed roundings, 8/16 deg4");
  run<2, 0, 4>("p2 0/16 poly (= round 1)");
  run<2, 2, 4>("p2 2/16 poly deg4");
  run<2, 3, 4>("p2 3/16 poly deg4");
  run<2, 4, 4>("p2 4/16 poly deg4");
  run<2, 5, 4>("p2 5/16 poly deg4");
  run<2, 6, 4>("p2 6/16 poly deg4");
  run<2, 4, 5>("p2 4/16 poly deg5");
  run<4, 0, 4>("p2 packed roundings, 0/16 poly");
  run<4, 3, 4>("p2 packed roundings, 3/16 deg4");
  run<4, 4, 4>("p2 packed roundings, 4/16 deg4");
  run<4, 6, 4>("p2 packed roundings, 6/16 deg4");
  return 0;
}
