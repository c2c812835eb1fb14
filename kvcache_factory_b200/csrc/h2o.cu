// H2O scoring (reference H2OKVCluster.update_kv, pyramidkv_utils.py:544-554): EVERY query row scores
// every key (non-causal except the bottom-right w x w block), probabilities are rounded to the model
// dtype and column-summed over all L rows.  The reference materialises [B,H,L,L] (64 GiB per sample at
// 32K); here the matrix only ever exists as 128 x 128 tiles in tensor memory.
//
//   pass 1  CTA = (batch, query head, block of 128 query rows): Q block stationary in shared memory,
//           ALL key tiles streamed by TMA, D = Q.K^T, thread <-> query row -> exact row max / sum-exp,
//           c_row = -M*log2e - log2 Z written to a [B,Hq,L] fp32 buffer.
//   pass 2  CTA = (batch, KV head, tile of 128 keys): K tile stationary, the g query heads' row blocks
//           streamed by TMA, D = K.Q^T, thread <-> key: p = exp2(l*log2e + c_row) -> dtype, summed over
//           rows per head in registers/shared memory (deterministic, no atomics), rounded once.
// Same rounding points as the window kernels (GEMM -> dtype, /sqrt(d) -> dtype, softmax fp32 -> dtype,
// column sum fp32 -> dtype); same warp specialisation (TMA warp, MMA warp, 8 epilogue warps, 4 TMEM
// accumulator buffers).  This path is exp-throughput bound (2*L^2 exps per head), not HBM bound.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_helpers.cuh"

namespace kvc {
namespace h2o {

constexpr int kTile = 128;
constexpr int kD = 128;
constexpr int kHalfBytes = kTile * 64 * 2;
constexpr int kTileBytes = 2 * kHalfBytes;
constexpr int kNumBuf = 4;
constexpr int kStages = 5;
constexpr int kEpiWarps = 16;  // 4 warpgroups; warpgroup g owns columns [32g, 32g+32) of every accumulator tile (as scores_fused.cu)
constexpr int kNumWG = kEpiWarps / 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kMaxGroups = 8;

struct Params {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  int B, Hq, Hkv, L, w, groups;
  int tiles;       // ceil(L / 128)
  float scale;
  float* cbuf;     // [B*Hq][tiles*128]  c_row, -inf for rows >= L
  uint16_t* out;   // [B,Hq,L-w]
};

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 22)) {
      printf("kvc h2o: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t instr_desc(int fmt) {  // M = N = 128
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(128 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

struct Smem {
  static constexpr int kStatBytes = kTileBytes;                 // stationary operand
  static constexpr int kRingBytes = kStages * kTileBytes;
  static constexpr int kBarOff = kStatBytes + kRingBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kNumBuf + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kAuxOff = (kTmemPtrOff + 4 + 15) & ~15;  // pass 1: float2[kNumWG][128]; pass 2: float[kNumWG][8][128]
  static constexpr int kTotal = kAuxOff + kNumWG * kMaxGroups * 128 * 4;
  static constexpr int kAlloc = kTotal + 1024;
};

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

constexpr float kTau = 6.0f;  // lazy running maximum, see scores_fused.cu

// Pass 1, one chunk of 32 keys of this thread's query row: the reference's two roundings (GEMM result
// -> dtype, * 1/sqrt(d) in fp32 -> dtype), lazy max, Z += sum exp(x - m_ref).
template <typename T, bool MASK>
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
  if (MASK) {
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
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], make_float2(fast_exp2(a.x), fast_exp2(a.y)));
  }
}

// Pass 2, one chunk of 32 query rows of this thread's key: p = exp2(x*log2e + c_row) rounded to dtype, summed in fp32.
template <typename T, int POLY>
__device__ __forceinline__ float p2_chunk(const uint32_t (&v)[32], float scale, const float4* __restrict__ cg) {
  const float2 scale2 = make_float2(scale, scale);
  const float2 l2e2 = make_float2(kLog2e, kLog2e);
  float2 acc[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
  for (int q4 = 0; q4 < 8; ++q4) {
    const float4 c4 = __ldg(cg + q4);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int i = 2 * q4 + h;
      float2 f = __fmul2_rn(make_float2(T::rnd(__uint_as_float(v[2 * i])), T::rnd(__uint_as_float(v[2 * i + 1]))), scale2);
      f = __ffma2_rn(make_float2(T::rnd(f.x), T::rnd(f.y)), l2e2, h == 0 ? make_float2(c4.x, c4.y) : make_float2(c4.z, c4.w));
      const float2 e = i < POLY ? tcx::exp2_fma2(f) : make_float2(fast_exp2(f.x), fast_exp2(f.y));   // FMA-pipe exp2 for POLY of 16 pairs
      acc[h] = __fadd2_rn(acc[h], make_float2(T::rnd(e.x), T::rnd(e.y)));
    }
  }
  return (acc[0].x + acc[0].y) + (acc[1].x + acc[1].y);
}

template <typename T, bool kPass2, int POLY>
__global__ void __launch_bounds__(kThreads, 1)
h2o_kernel(const __grid_constant__ CUtensorMap tmap_k, const __grid_constant__ CUtensorMap tmap_q, const Params p) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - raw_addr);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nscore = p.L - p.w;

  // pass 1: blockIdx.x = query row block, blockIdx.y = b*Hq + hq; pass 2: blockIdx.x = key tile, blockIdx.y = b*Hkv + kvh
  const int blk = blockIdx.x;
  const int b = kPass2 ? blockIdx.y / p.Hkv : blockIdx.y / p.Hq;
  const int hq = kPass2 ? 0 : blockIdx.y % p.Hq;
  const int kvh = kPass2 ? blockIdx.y % p.Hkv : hq / p.groups;
  const int nitems = kPass2 ? p.groups * p.tiles : p.tiles;  // streamed tiles

  const uint32_t stat_s = base, ring_s = base + Smem::kStatBytes, bar_s = base + Smem::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (kStages + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * kStages + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * kStages + kNumBuf + i); };
  const uint32_t sbar = bar_s + 8u * (2 * kStages + 2 * kNumBuf);  // stationary operand staged
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + Smem::kTmemPtrOff);
  float* aux = reinterpret_cast<float*>(smem + Smem::kAuxOff);

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(kfull(s), 1); mbar_init(kempty(s), 1); }
    for (int i = 0; i < kNumBuf; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), kEpiWarps); }
    mbar_init(sbar, kPass2 ? 1 : kEpiWarps * 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if constexpr (kPass2) {
    for (int i = tid; i < kNumWG * kMaxGroups * 128; i += kThreads) aux[i] = 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp >= 2) {
    if constexpr (!kPass2) {
      // stationary Q block: rows blk*128 .. +127 of head hq, canonical K-major SWIZZLE_128B layout
      const int et = tid - 64;
#pragma unroll
      for (int it = 0; it < 128 * 16 / (kEpiWarps * 32); ++it) {
        const int e = et + it * (kEpiWarps * 32);
        const int r = e >> 4, c16 = e & 15;
        const int row = blk * kTile + r;
        const bool ok = row < p.L;
        const uint16_t* src = p.q + b * p.q_sb + (int64_t)hq * p.q_sh + (int64_t)(ok ? row : 0) * p.q_sl + c16 * 8;
        const uint32_t dst = stat_s + (c16 >> 3) * kHalfBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
        const int sz = ok ? 16 : 0;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      asm volatile("cp.async.wait_all;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      mbar_arrive(sbar);
    }
  }

  if (warp == 0) {
    // TMA producer: warp-uniform loop, one elected lane issues (see scores_fused.cu)
    if constexpr (kPass2) {  // stationary K tile
      if (elect_one()) {
        mbar_expect_tx(sbar, kTileBytes);
        tma_load_4d(stat_s, &tmap_k, sbar, 0, blk * kTile, kvh, b);
        tma_load_4d(stat_s + kHalfBytes, &tmap_k, sbar, 64, blk * kTile, kvh, b);
      }
      __syncwarp();
    }
    int s = 0, ph = 1, hl = 0, rbk = 0;
    for (int t = 0; t < nitems; ++t) {
      mbar_wait(kempty(s), ph);
      if (elect_one()) {
        mbar_expect_tx(kfull(s), kTileBytes);
        if constexpr (kPass2) {  // heads interleaved: consecutive tiles are the g heads of one query row block
          tma_load_4d(ring_s + s * kTileBytes, &tmap_q, kfull(s), 0, rbk * kTile, kvh * p.groups + hl, b);
          tma_load_4d(ring_s + s * kTileBytes + kHalfBytes, &tmap_q, kfull(s), 64, rbk * kTile, kvh * p.groups + hl, b);
        } else {
          tma_load_4d(ring_s + s * kTileBytes, &tmap_k, kfull(s), 0, t * kTile, kvh, b);
          tma_load_4d(ring_s + s * kTileBytes + kHalfBytes, &tmap_k, kfull(s), 64, t * kTile, kvh, b);
        }
      }
      __syncwarp();
      if (++s == kStages) { s = 0; ph ^= 1; }
      if (++hl == p.groups) { hl = 0; ++rbk; }
    }
  } else if (warp == 1) {
    // MMA issuer: warp-uniform loop, elected lane issues; descriptors stay in uniform registers
    const uint32_t idesc = instr_desc(T::kDtype == KVC_BF16 ? 1 : 0);
    mbar_wait(sbar, 0);
    tc_fence_after();
    const uint64_t ad0 = smem_desc(stat_s);
    int s = 0, ph = 0, buf = 0, bph = 1;
    for (int t = 0; t < nitems; ++t) {
      mbar_wait(tempty(buf), bph);
      mbar_wait(kfull(s), ph);
      tc_fence_after();
      const uint64_t bd0 = smem_desc(ring_s + s * kTileBytes);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t off = (uint64_t)(((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4);
          // A (M = TMEM lanes) is always the stationary tile: query rows in pass 1, keys in pass 2
          tc_mma(tmem_base + buf * 128, ad0 + off, bd0 + off, idesc, k > 0);
        }
        tc_commit(tfull(buf));
        tc_commit(kempty(s));
      }
      __syncwarp();
      if (++s == kStages) { s = 0; ph ^= 1; }
      if (++buf == kNumBuf) { buf = 0; bph ^= 1; }
    }
  } else {
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r_loc = quad * 32 + lane;
    const float scale = p.scale;
    const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + wg * 32;
    float m_run = -1e30f;
    float2 zacc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    int hl = 0, rbk = 0;

    for (int j = 0; j < nitems; ++j) {
      const int buf = j % kNumBuf;
      mbar_wait(tfull(buf), (j / kNumBuf) & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld32(tcol + buf * 128, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));

      if constexpr (!kPass2) {
        // thread <-> query row `row`; this warpgroup's columns are keys key0 .. key0+31
        const int row = blk * kTile + r_loc, key0 = j * kTile + wg * 32;
        if (key0 + 32 > nscore) {  // chunk touches the w x w causal block or runs past L
          const int last_key = min(p.L - 1, row >= nscore ? row : p.L - 1);
          tcx::p1_chunk<T, true, POLY>(v, scale, m_run, zacc, key0, last_key);
        } else {
          tcx::p1_chunk<T, false, POLY>(v, scale, m_run, zacc, 0, 0);
        }
      } else {
        // thread <-> key blk*128 + r_loc; this warpgroup's columns are query rows 32*wg.. of (head hl, row block rbk)
        const float4* cg = reinterpret_cast<const float4*>(p.cbuf + ((int64_t)b * p.Hq + kvh * p.groups + hl) * (p.tiles * kTile) +
                                                            (int64_t)rbk * kTile + wg * 32);
        float* slot = aux + (wg * kMaxGroups + hl) * 128 + r_loc;  // only this thread touches this slot
        *slot += p2_chunk<T, POLY>(v, scale, cg);
        if (++hl == p.groups) { hl = 0; ++rbk; }
      }
    }
    if constexpr (!kPass2) {
      const float2 z2 = __fadd2_rn(__fadd2_rn(zacc[0], zacc[1]), __fadd2_rn(zacc[2], zacc[3]));
      reinterpret_cast<float2*>(aux)[wg * 128 + r_loc] = make_float2(m_run, z2.x + z2.y);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (!kPass2) {
    for (int r = tid; r < 128; r += kThreads) {
      float mn = -1e30f, z = 0.f;
#pragma unroll
      for (int g = 0; g < kNumWG; ++g) mn = fmaxf(mn, reinterpret_cast<float2*>(aux)[g * 128 + r].x);
#pragma unroll
      for (int g = 0; g < kNumWG; ++g) {
        const float2 a = reinterpret_cast<float2*>(aux)[g * 128 + r];
        z += a.y * fast_exp2((a.x - mn) * kLog2e);
      }
      const int row = blk * kTile + r;
      p.cbuf[((int64_t)b * p.Hq + hq) * (p.tiles * kTile) + row] = row < p.L ? (-mn * kLog2e - log2f(z)) : -INFINITY;
    }
  } else {
    for (int e = tid; e < p.groups * 128; e += kThreads) {
      const int hl = e >> 7, r = e & 127, key = blk * kTile + r;
      if (key < nscore) {
        float s = 0.f;
#pragma unroll
        for (int g = 0; g < kNumWG; ++g) s += aux[(g * kMaxGroups + hl) * 128 + r];  // fixed order: deterministic
        p.out[((int64_t)b * p.Hq + kvh * p.groups + hl) * nscore + key] = T::round1(s);
      }
    }
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
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

// [B,H,L,d] tensor with element strides as a 4-D TMA map {d, L, H, B}, box = 64 columns x 128 rows.
static int make_map(CUtensorMap* tmap, const void* ptr, int64_t sb, int64_t sh, int64_t sl, int B, int H, int L, int d,
                    int dtype) {
  EncodeTiledFn enc = encode_fn();
  KVC_REQUIRE(enc != nullptr, KVC_ERR_DEVICE, "cuTensorMapEncodeTiled is not available from this driver");
  const cuuint64_t gdim[4] = {(cuuint64_t)d, (cuuint64_t)L, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)sl * 2, (cuuint64_t)sh * 2, (cuuint64_t)sb * 2};
  if (H == 1) gstr[1] = gstr[0] * (cuuint64_t)L;
  if (B == 1) gstr[2] = gstr[0] * (cuuint64_t)L * (cuuint64_t)H;
  const cuuint32_t box[4] = {64, (cuuint32_t)kTile, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(tmap, dtype == KVC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                          const_cast<void*>(ptr), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVC_REQUIRE(cr == CUDA_SUCCESS, KVC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
  return 0;
}

template <typename T, int POLY>
static int launch(const CUtensorMap& mk, const CUtensorMap& mq, const Params& sp, cudaStream_t stream) {
  auto k1 = h2o_kernel<T, false, POLY>;
  auto k2 = h2o_kernel<T, true, POLY>;
  KVC_SET_MAX_SMEM(k1, Smem::kAlloc);
  KVC_SET_MAX_SMEM(k2, Smem::kAlloc);
  k1<<<dim3(sp.tiles, sp.B * sp.Hq), kThreads, Smem::kAlloc, stream>>>(mk, mq, sp);
  if (int rc = launch_status("h2o_kernel<stats>")) return rc;
  const int ktiles = (sp.L - sp.w + kTile - 1) / kTile;
  k2<<<dim3(ktiles, sp.B * sp.Hkv), kThreads, Smem::kAlloc, stream>>>(mk, mq, sp);
  return launch_status("h2o_kernel<scores>");
}

}  // namespace h2o

size_t h2o_scores_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  (void)Hkv; (void)d; (void)w;
  const size_t tiles = (size_t)(L + h2o::kTile - 1) / h2o::kTile;
  return ((size_t)B * Hq * tiles * h2o::kTile * sizeof(float) + 255) & ~(size_t)255;
}

int h2o_scores(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sh,
               int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, void* out_scores, void* workspace,
               size_t workspace_bytes, cudaStream_t stream) {
  using namespace h2o;
  KVC_REQUIRE(d == 128, KVC_ERR_UNSUPPORTED, "H2O scoring kernel implements head_dim 128 (got %d)", d);
  KVC_REQUIRE(Hq / Hkv <= kMaxGroups, KVC_ERR_UNSUPPORTED, "H2O scoring kernel: at most %d query heads per KV head", kMaxGroups);
  KVC_REQUIRE(((q_sb | q_sh | q_sl | k_sb | k_sh | k_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0 &&
                  (reinterpret_cast<uintptr_t>(k) & 15) == 0,
              KVC_ERR_UNSUPPORTED, "H2O scoring kernel: query/key rows must be 16-byte aligned (TMA)");
  KVC_REQUIRE(workspace_bytes >= h2o_scores_workspace_bytes(B, Hq, Hkv, L, d, w), KVC_ERR_WORKSPACE,
              "h2o_scores workspace: have %zu need %zu", workspace_bytes, h2o_scores_workspace_bytes(B, Hq, Hkv, L, d, w));
  CUtensorMap mk, mq;
  if (int rc = make_map(&mk, k, k_sb, k_sh, k_sl, B, Hkv, L, d, dtype)) return rc;
  if (int rc = make_map(&mq, q, q_sb, q_sh, q_sl, B, Hq, L, d, dtype)) return rc;
  Params sp{};
  sp.q = (const uint16_t*)q; sp.q_sb = q_sb; sp.q_sh = q_sh; sp.q_sl = q_sl;
  sp.B = B; sp.Hq = Hq; sp.Hkv = Hkv; sp.L = L; sp.w = w; sp.groups = Hq / Hkv;
  sp.tiles = (L + kTile - 1) / kTile;
  sp.scale = 1.0f / sqrtf((float)d);
  sp.cbuf = (float*)workspace;
  sp.out = (uint16_t*)out_scores;
  // KVC_H2O_POLY=1: 4 of every 16 pairs of exponentials on the FMA pipe instead of MUFU (tc_helpers.cuh).  Measured
  // SLOWER (L = 16384, B = 2: 11.26 ms with 0 of 16, 11.61 with 2, 12.7 with 4, 13.1 with 6): the loop is bound by
  // instruction issue (ncu: issue 67 %, XU 68 %, ALU 61 %), and a polynomial exp2 costs 5.5 issue slots against 1.
  static const bool poly = [] { const char* e = getenv("KVC_H2O_POLY"); return e && atoi(e) > 0; }();
  if (poly) return dtype == KVC_BF16 ? launch<Bf16, 4>(mk, mq, sp, stream) : launch<Fp16, 4>(mk, mq, sp, stream);
  return dtype == KVC_BF16 ? launch<Bf16, 0>(mk, mq, sp, stream) : launch<Fp16, 0>(mk, mq, sp, stream);
}

}  // namespace kvc
