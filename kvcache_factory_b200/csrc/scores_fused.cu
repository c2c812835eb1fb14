// Window scoring, single-launch version of scores_tc.cu: both softmax passes in ONE persistent kernel
// with a grid-wide barrier in between.
//
//   every CTA:  pass 1 over its chunk of keys (stats)  ->  publish (m, Z) partials  ->  grid barrier
//               ->  fold all chunks' partials into c_row  ->  pass 2 over the same chunk (scores)
//
// Only the 8 epilogue warps wait at the barrier.  The TMA producer and the MMA issuer run straight
// on into pass 2 (K tiles and Q.K^T do not depend on the statistics), so when the barrier opens the
// 4 TMEM accumulator buffers already hold pass-2 tiles and the K ring is full: the barrier costs the
// skew between CTAs, not a pipeline refill.  Pass 2 walks the chunk backwards, i.e. starts with the
// tiles that were read last (still in L2).  Versus two launches this saves a kernel launch + ramp,
// one Q staging / TMEM allocation / stats re-read, and the drain between the kernels.
//
// Requires all CTAs to be co-resident (grid <= #SMs, one CTA per SM: the grid is sized that way).
// Math, rounding points and operand roles are identical to scores_tc.cu.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "common.cuh"

namespace kvc {
namespace tcf {

constexpr int kTile = 128;
constexpr int kD = 128;
constexpr int kHalfBytes = kTile * 64 * 2;
constexpr int kTileBytes = 2 * kHalfBytes;
constexpr int kNumBuf = 4;
// Four epilogue warpgroups; warpgroup g owns columns [32g, 32g+32) of EVERY accumulator tile.  A tile's
// TMEM buffer is therefore drained by all 16 warps at once and handed back after one 32-column load
// each, which keeps the MMA issuer up to three tiles ahead of the slowest warp (with one tile per
// warpgroup the buffers were held ~4x longer and the epilogue starved ~30% of the time); 4 warps per
// scheduler cover the TMEM-load and MUFU latencies.
constexpr int kEpiWarps = 16;
constexpr int kNumWG = kEpiWarps / 4;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 32 * (2 + kEpiWarps);

struct Params {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  int B, Hq, Hkv, L, w, groups;
  int chunk_tiles, nchunks;
  float scale;
  int reduce_mean;
  int rbs;              // half-footprint kernel: 128-row blocks (= units) per KV head
  float2* stats;        // [units][nchunks][RB*128]
  uint16_t* out;        // [B,Hq,L-w]
  unsigned long long* tl;  // developer timeline entry (TRACE builds), else nullptr
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
      printf("kvc scores_fused: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
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
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ uint32_t instr_desc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// Optional timeline (developer tool, `make TRACE=1`): 16 globaltimer stamps per CTA, read back by
// kvc_debug_ftrace().  Compiled out of the shipped library.
#ifdef KVC_TRACE
__device__ long long g_trace[160 * 16];
__device__ unsigned g_tile_trace[4][64][8];  // CTAs 0, 1, 70, 143: per tile SM-clock stamps {tma issue, kfull seen, tempty seen, mma committed, epi got tile, epi loaded, epi done}
__shared__ unsigned s_tile_trace[64][8];
__device__ __forceinline__ void tile_stamp(int j, int slot) {
  if (j < 64) s_tile_trace[j][slot] = (unsigned)clock64();
}
#define KVC_TILE_STAMP(cond, j, slot) do { if (cond) tile_stamp(j, slot); } while (0)
__device__ __forceinline__ void stamp(int slot) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_trace[(blockIdx.y * gridDim.x + blockIdx.x) * 16 + slot] = t;
}
#define KVC_STAMP(cond, slot) do { if (cond) stamp(slot); } while (0)
#else
#define KVC_STAMP(cond, slot) do { } while (0)
#define KVC_TILE_STAMP(cond, j, slot) do { } while (0)
#endif

// Grid-barrier state (per device, zero at module load, self-resetting).  One barrier kernel runs at a
// time on a device: the cooperative launch admits a grid only when all of its CTAs fit, and two such
// grids (one CTA per SM each) never fit together.
__device__ unsigned g_bar_count = 0, g_bar_gen = 0;

template <int RB>
struct Smem {
  static constexpr int kStages = RB == 1 ? 5 : 4;
  static constexpr int kQBytes = RB * kTileBytes;
  static constexpr int kKBytes = kStages * kTileBytes;
  static constexpr int kBarOff = kQBytes + kKBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kNumBuf + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kCrowOff = (kTmemPtrOff + 4 + 15) & ~15;        // float[RB*128]
  static constexpr int kXstatOff = kCrowOff + RB * 128 * 4;            // float2[RB][kNumWG][128]
  static constexpr int kTotal = kXstatOff + RB * kNumWG * 128 * 8;
  static constexpr int kAlloc = kTotal + 1024;
};

// Lazy running maximum of pass 1: the exponentials of a row are taken relative to a reference m_ref
// that is only raised when a 32-key chunk exceeds it by more than kTau (a warp-uniform, rare branch);
// (m_ref, Z) still represents Z * e^m_ref exactly, so the folded statistics are unchanged while the
// per-chunk rescale (one MUFU + a dependent multiply per row per chunk) disappears from the hot loop.
constexpr float kTau = 6.0f;

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
    zacc[i & 3] = __fadd2_rn(zacc[i & 3], make_float2(fast_exp2(a.x), fast_exp2(a.y)));
  }
}

// Pass 2, one chunk of 32 stacked query rows of this thread's key: the same two roundings,
// p = exp2(x*log2e + c_row) rounded to dtype (softmax output), summed per head in fp32.
template <typename T, int W>
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
      const float2 e = make_float2(T::rnd(fast_exp2(f.x)), T::rnd(fast_exp2(f.y)));
      acc[(2 * i) / W] = __fadd2_rn(acc[(2 * i) / W], e);
    }
  }
}

template <typename T, int RB, int W>
__global__ void __launch_bounds__(kThreads, 1)
window_score_fused_kernel(const __grid_constant__ CUtensorMap tmap_k, const Params p) {
  using SM = Smem<RB>;
  constexpr int S = SM::kStages;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x, unit = blockIdx.y;
  const int b = unit / p.Hkv, kvh = unit % p.Hkv;
  const int nscore = p.L - p.w;
  const int tile_begin = chunk * p.chunk_tiles;
  const int n1 = max(0, min(tile_begin + p.chunk_tiles, (p.L + kTile - 1) / kTile) - tile_begin);    // pass-1 tiles
  const int n2 = max(0, min(tile_begin + p.chunk_tiles, (nscore + kTile - 1) / kTile) - tile_begin); // pass-2 tiles
  const int nq = RB == 1 ? p.groups * p.w : 128;
  // item it < n1: pass-1 tile tile_begin+it; item it >= n1: pass-2 tile, walking the chunk backwards
  auto tile_of = [&](int it) { return it < n1 ? tile_begin + it : tile_begin + (n2 - 1 - (it - n1)); };

  const uint32_t q_s = base, k_s = base + SM::kQBytes, bar_s = base + SM::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (S + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * S + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * S + kNumBuf + i); };
  const uint32_t qbar = bar_s + 8u * (2 * S + 2 * kNumBuf);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::kTmemPtrOff);
  float* crow = reinterpret_cast<float*>(smem + SM::kCrowOff);
  float2* xstat = reinterpret_cast<float2*>(smem + SM::kXstatOff);

  KVC_STAMP(tid == 0, 0);
  tl_start(p.tl);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(kfull(s), 1); mbar_init(kempty(s), 1); }
    for (int i = 0; i < kNumBuf; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), kEpiWarps); }
    mbar_init(qbar, kEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;
  // Programmatic dependent launch: everything above (barrier init, TMEM allocation) may overlap the tail
  // of the previous kernel in the stream; nothing below may start before that kernel has completed and
  // its writes are visible (inputs, the statistics workspace, the grid-barrier words).  Letting our own
  // dependents launch right away is safe for the same reason: they wait for our completion in turn.
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== TMA producer: one stream of n1 + n2 tiles =====
    // The whole warp runs the loop (warp-uniform control flow keeps addresses and descriptors in uniform
    // registers); one elected lane issues.
    int s = 0, ph = 1;
    for (int it = 0; it < n1 + n2; ++it) {
      mbar_wait(kempty(s), ph);
      if (elect_one()) {
        mbar_expect_tx(kfull(s), kTileBytes);
        KVC_TILE_STAMP(true, it, 0);
        const int key0 = tile_of(it) * kTile;
        tma_load_4d(k_s + s * kTileBytes, &tmap_k, kfull(s), 0, key0, kvh, b);
        tma_load_4d(k_s + s * kTileBytes + kHalfBytes, &tmap_k, kfull(s), 64, key0, kvh, b);
        KVC_STAMP(it == n1 - 1, 12);       // last pass-1 tile requested
        KVC_STAMP(it == n1 + n2 - 1, 13);  // last pass-2 tile requested
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: pass 1 D = Q.K^T (lanes = query rows), pass 2 D = K.Q^T (lanes = keys) =====
    // Warp-uniform loop, elected lane issues: with the loop inside `if (lane == 0)` every descriptor
    // went through R2UR and one tile took ~950 cycles to issue (more than its 512 tensor-pipe cycles).
    const int fmt = T::kDtype == KVC_BF16 ? 1 : 0;
    const uint32_t idesc1 = instr_desc(fmt, 128, 128), idesc2 = instr_desc(fmt, 128, nq);
    mbar_wait(qbar, 0);
    tc_fence_after();
    int s = 0, ph = 0, buf = 0, bph = 1;
    for (int it = 0; it < n1 + n2; ++it) {
      const bool second = it >= n1;
      const uint32_t idesc = second ? idesc2 : idesc1;
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        mbar_wait(tempty(buf), bph);
        if (rb == 0) mbar_wait(kfull(s), ph);
        tc_fence_after();
        KVC_TILE_STAMP(lane == 0, it, 2);
        const uint64_t kd0 = smem_desc(k_s + s * kTileBytes), qd0 = smem_desc(q_s + rb * kTileBytes);
        const uint64_t ad0 = second ? kd0 : qd0, bd0 = second ? qd0 : kd0;
        if (elect_one()) {
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint64_t off = (uint64_t)(((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4);  // start-address field is in 16-byte units
            tc_mma(tmem_base + buf * 128, ad0 + off, bd0 + off, idesc, k > 0);
          }
          tc_commit(tfull(buf));
          if (rb == RB - 1) tc_commit(kempty(s));
        }
        __syncwarp();
        KVC_TILE_STAMP(lane == 0, it, 3);
        if (++buf == kNumBuf) { buf = 0; bph ^= 1; }
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    // ===== epilogue warps =====
    const int et = tid - 64;
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r_loc = quad * 32 + lane;
    float scale = p.scale;
    unsigned my_gen = 0;
    if (et == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(my_gen) : "l"(&g_bar_gen) : "memory");
    // stage the Q block(s): canonical K-major SWIZZLE_128B layout, all chunks in flight
#pragma unroll
    for (int it = 0; it < (RB * 128 * 16 + kEpiThreads - 1) / kEpiThreads; ++it) {
      const int e = et + it * kEpiThreads;
      if (e >= RB * 128 * 16) break;
      const int rg = e >> 4, c16 = e & 15;
      const int rb = rg >> 7, r = rg & 127;
      const int hl = rg / W, i = rg % W;
      const bool ok = hl < p.groups;
      const uint16_t* src = p.q + b * p.q_sb + (int64_t)(kvh * p.groups + (ok ? hl : 0)) * p.q_sh +
                            (int64_t)(nscore + i) * p.q_sl + c16 * 8;
      const uint32_t dst = q_s + rb * kTileBytes + (c16 >> 3) * kHalfBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(qbar);
    KVC_STAMP(et == 0, 1);

    // ---------------- pass 1: thread <-> query row, columns are keys ----------------
    float m_run[RB];
    float2 zacc[RB][4];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      m_run[rb] = -1e30f;
#pragma unroll
      for (int a = 0; a < 4; ++a) zacc[rb][a] = make_float2(0.f, 0.f);
    }
    // (Two software-pipelined variants were measured slower, 36.9 / 42 us vs 34.4 us per launch: the TMEM
    // load of tile j+1 issued before the math of tile j into a second register buffer, and issued
    // between the convert and exp halves of tile j into the registers the convert half had just freed.)
    const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + wg * 32;
    // small windows (g*w < 128): the accumulator rows beyond g*w are zero padding -- warps that own only
    // padding rows skip the load and the 32 exps per tile (w = 8 with 4 groups: 3 of 4 warps per warpgroup)
    bool dead_rows[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) dead_rows[rb] = __all_sync(0xffffffffu, (rb * 128 + r_loc) / W >= p.groups);
    int j = 0;
    for (; j < n1 * RB; ++j) {
      const int it = j / RB, rb = j % RB, buf = j % kNumBuf;
      const int key0 = tile_of(it) * kTile + wg * 32;
      mbar_wait(tfull(buf), (j / kNumBuf) & 1);
      tc_fence_after();
      KVC_TILE_STAMP(et == 0, j, 4);
      if (dead_rows[rb]) {  // all 32 rows of this warp are padding (g*w < 128): nothing to add up
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
        continue;
      }
      uint32_t v[32];
      tmem_ld32(tcol + buf * 128, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));
      KVC_TILE_STAMP(et == 0, j, 5);
      if (key0 + 32 > nscore) {  // chunk touches the w x w causal block or the ragged end
        const int rg = rb * 128 + r_loc;
        const int last_key = min(p.L - 1, rg / W < p.groups ? nscore + rg % W : p.L - 1);
        p1_chunk<T, true>(v, scale, m_run[rb], zacc[rb], key0, last_key);
      } else {
        p1_chunk<T, false>(v, scale, m_run[rb], zacc[rb], 0, 0);
      }
      KVC_TILE_STAMP(et == 0 && zacc[rb][0].x >= 0.f, j, 6);
    }
    KVC_STAMP(et == 0, 2);  // pass 1 done
    float z_run[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const float2 z2 = __fadd2_rn(__fadd2_rn(zacc[rb][0], zacc[rb][1]), __fadd2_rn(zacc[rb][2], zacc[rb][3]));
      z_run[rb] = z2.x + z2.y;
    }

    // ---------------- publish partials, grid barrier, fold all chunks into c_row ----------------
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) xstat[(rb * kNumWG + wg) * 128 + r_loc] = make_float2(m_run[rb], z_run[rb]);
    epi_sync();
    for (int r = et; r < RB * 128; r += kEpiThreads) {
      const int rb = r >> 7, rl = r & 127;
      float mn = -1e30f, z = 0.f;
#pragma unroll
      for (int g = 0; g < kNumWG; ++g) mn = fmaxf(mn, xstat[(rb * kNumWG + g) * 128 + rl].x);
#pragma unroll
      for (int g = 0; g < kNumWG; ++g) {
        const float2 a = xstat[(rb * kNumWG + g) * 128 + rl];
        z += a.y * fast_exp2((a.x - mn) * kLog2e);
      }
      p.stats[((int64_t)unit * p.nchunks + chunk) * (RB * 128) + r] = make_float2(mn, z);
    }
    __threadfence();
    epi_sync();
    KVC_STAMP(et == 0, 5);  // partials published
    if (et == 0) {
      // sense-reversing grid barrier on library-owned words: the last CTA to arrive resets the count
      // and bumps the generation, so no memset node is needed between launches
      const unsigned total = gridDim.x * gridDim.y;
      const unsigned old = atomicAdd(&g_bar_count, 1u);
      if (old == total - 1) {
        g_bar_count = 0u;
        __threadfence();
        atomicAdd(&g_bar_gen, 1u);
      } else {
        unsigned gen = my_gen;
        for (unsigned spin = 0; gen == my_gen; ++spin) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&g_bar_gen) : "memory");
          if (gen == my_gen) {
            __nanosleep(32);
            if (spin > (1u << 24)) {
              printf("kvc scores_fused: grid barrier timed out (block %d,%d)\n", blockIdx.x, blockIdx.y);
              __trap();
            }
          }
        }
      }
    }
    epi_sync();
    KVC_STAMP(et == 0, 6);  // grid barrier passed
    {
      // fold every chunk's (m, Z): all 512 threads take part -- thread = (row, part), part q folds chunks
      // q, q+nparts, ... (loads in one or two batches instead of nchunks/8 dependent round trips per row),
      // the parts meet in shared memory
      constexpr int kRows = RB * 128, kParts = kEpiThreads / kRows;
      const int r = et % kRows, q = et / kRows;
      const float2* st = p.stats + (int64_t)unit * p.nchunks * kRows + r;
      float M = -1e30f, Z = 0.f;
      for (int c0 = q; c0 < p.nchunks; c0 += 8 * kParts) {
        float2 sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sv[u] = make_float2(-1e30f, 0.f);
          if (c0 + u * kParts < p.nchunks) {
            const float2* a = st + (int64_t)(c0 + u * kParts) * kRows;
            asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(sv[u].x), "=f"(sv[u].y) : "l"(a) : "memory");  // written by other SMs: bypass L1
          }
        }
        float mx = M;
#pragma unroll
        for (int u = 0; u < 8; ++u) mx = fmaxf(mx, sv[u].x);
        Z *= fast_exp2((M - mx) * kLog2e);
#pragma unroll
        for (int u = 0; u < 8; ++u) Z += sv[u].y * fast_exp2((sv[u].x - mx) * kLog2e);
        M = mx;
      }
      xstat[q * kRows + r] = make_float2(M, Z);
      epi_sync();
      if (et < kRows) {
        float mx = -1e30f, z = 0.f;
#pragma unroll
        for (int g = 0; g < kParts; ++g) mx = fmaxf(mx, xstat[g * kRows + et].x);
#pragma unroll
        for (int g = 0; g < kParts; ++g) {
          const float2 a = xstat[g * kRows + et];
          z += a.y * fast_exp2((a.x - mx) * kLog2e);
        }
        crow[et] = (et / W < p.groups) ? (-mx * kLog2e - log2f(z)) : -INFINITY;
      }
    }
    epi_sync();

    KVC_STAMP(et == 0, 7);  // c_row ready
    // ---------------- pass 2: thread <-> key, columns are the stacked query rows ----------------
    constexpr int kHeadsPerChunk = 32 / W;
    for (; j < (n1 + n2) * RB; ++j) {
      const int it = j / RB, rb = j % RB, buf = j % kNumBuf;
      const int key = tile_of(it) * kTile + r_loc;
      mbar_wait(tfull(buf), (j / kNumBuf) & 1);
      tc_fence_after();
      KVC_TILE_STAMP(et == 0, j, 4);
      if ((rb * 128 + wg * 32) / W >= p.groups) {  // this warpgroup's 32 columns are beyond the last head
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
        continue;
      }
      uint32_t v[32];
      tmem_ld32(tcol + buf * 128, v);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));
      KVC_TILE_STAMP(et == 0, j, 5);
      float2 acc[kHeadsPerChunk];
#pragma unroll
      for (int h = 0; h < kHeadsPerChunk; ++h) acc[h] = make_float2(0.f, 0.f);
      p2_chunk<T, W>(v, scale, crow + rb * 128 + wg * 32, acc);
      if (key < nscore) {
#pragma unroll
        for (int h = 0; h < kHeadsPerChunk; ++h) {
          const int head = (rb * 128 + wg * 32) / W + h;
          if (head < p.groups) {
            float s = acc[h].x + acc[h].y;
            if (p.reduce_mean) s = div_exact(s, (float)p.w, 1.0f / (float)p.w);
            p.out[((int64_t)b * p.Hq + kvh * p.groups + head) * nscore + key] = T::round1(s);
          }
        }
      }
      KVC_TILE_STAMP(et == 0 && acc[0].x >= 0.f, j, 6);
    }
  }

  KVC_STAMP(tid == 64, 8);  // pass 2 done
  tc_fence_before();
  __syncthreads();
  KVC_STAMP(tid == 0, 11);
#ifdef KVC_TRACE
  {
    const int cta = blockIdx.y * gridDim.x + blockIdx.x;
    const int c = cta == 0 ? 0 : cta == 1 ? 1 : cta == 70 ? 2 : cta == 143 ? 3 : -1;
    if (c >= 0) for (int i = tid; i < 64 * 8; i += kThreads) g_tile_trace[c][i / 8][i % 8] = s_tile_trace[i / 8][i % 8];
  }
#endif
  tl_end(p.tl);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// =====================================================================================================
// Half-footprint variant: the same two passes in a CTA HALF the size (2 accumulator buffers = 256 TMEM
// columns, 2 K stages, ~100 KB of shared memory, half the register file) so that TWO CTAs --
// of two different launches -- share an SM.  A single launch leaves ~9 us of every ~34 us idle on every SM
// (prologue, grid-barrier skew, statistics fold, tail); with the scoring launches of consecutive layers on
// two streams those gaps are filled by the other layer's CTA.  Warpgroup g owns columns [g*128/NWG, (g+1)*128/NWG)
// of every tile.  Two launches are in flight at once, so the grid barrier uses one
// pair of library-owned, self-resetting words per STREAM.  RB == 1 only.
// NWG warpgroups share every tile: 2 (two 32-column chunks each, <= 96 registers) or 4 (one chunk each, <= 56
// registers: twice the warps per SM sub-partition to cover each other's tcgen05.ld and barrier latencies).
constexpr int kHNumBuf = 2;

struct HalfSmem {
  static constexpr int kStages = 2;
  static constexpr int kQBytes = kTileBytes;
  static constexpr int kKBytes = kStages * kTileBytes;
  static constexpr int kBarOff = kQBytes + kKBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kHNumBuf + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kCrowOff = (kTmemPtrOff + 4 + 15) & ~15;   // float[128]
  static constexpr int kXstatOff = kCrowOff + 128 * 4;            // float2[NWG <= 4][128]
  static constexpr int kTotal = kXstatOff + 4 * 128 * 8;
  static constexpr int kAlloc = kTotal + 1024;
};

constexpr int kHSlots = 16;
__device__ unsigned g_hbar_count[kHSlots], g_hbar_gen[kHSlots];   // zero at module load, self-resetting

template <int kN>
__device__ __forceinline__ void hepi_sync_n() { asm volatile("bar.sync 1, %0;" ::"n"(kN) : "memory"); }

template <typename T, int W, int NWG>
__global__ void __launch_bounds__(64 + 128 * NWG, 2)
window_score_half_kernel(const __grid_constant__ CUtensorMap tmap_k, const Params p, int slot) {
  using SM = HalfSmem;
  constexpr int kHEpiWarps = 4 * NWG, kHEpiThreads = 128 * NWG;
  constexpr int kCols = 128 / NWG, kChunks = kCols / 32;   // columns / 32-column chunks of a tile per warpgroup
  auto hepi_sync = [] { hepi_sync_n<kHEpiThreads>(); };
  constexpr int S = SM::kStages;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int chunk = blockIdx.x, unit = blockIdx.y;
  // a unit is one 128-row block of a (batch, KV head): KV heads with more than 128 stacked window rows (70B: 8 x 32)
  // are served as p.rbs units that read the same keys (the second one from L2)
  const int rbi = unit % p.rbs, kvu = unit / p.rbs;
  const int b = kvu / p.Hkv, kvh = kvu % p.Hkv;
  const int head0 = rbi * (128 / W);                       // first head of the KV group in this row block
  const int gl = min(128 / W, p.groups - head0);           // heads in this row block
  const int nscore = p.L - p.w;
  const int tile_begin = chunk * p.chunk_tiles;
  const int n1 = max(0, min(tile_begin + p.chunk_tiles, (p.L + kTile - 1) / kTile) - tile_begin);
  const int n2 = max(0, min(tile_begin + p.chunk_tiles, (nscore + kTile - 1) / kTile) - tile_begin);
  const int nq = gl * p.w;
  auto tile_of = [&](int it) { return it < n1 ? tile_begin + it : tile_begin + (n2 - 1 - (it - n1)); };

  const uint32_t q_s = base, k_s = base + SM::kQBytes, bar_s = base + SM::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (S + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * S + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * S + kHNumBuf + i); };
  const uint32_t qbar = bar_s + 8u * (2 * S + 2 * kHNumBuf);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::kTmemPtrOff);
  float* crow = reinterpret_cast<float*>(smem + SM::kCrowOff);
  float2* xstat = reinterpret_cast<float2*>(smem + SM::kXstatOff);

  tl_start(p.tl);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(kfull(s), 1); mbar_init(kempty(s), 1); }
    for (int i = 0; i < kHNumBuf; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), kHEpiWarps); }
    mbar_init(qbar, kHEpiThreads);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(256u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    int s = 0, ph = 1;
    for (int it = 0; it < n1 + n2; ++it) {
      mbar_wait(kempty(s), ph);
      if (elect_one()) {
        mbar_expect_tx(kfull(s), kTileBytes);
        const int key0 = tile_of(it) * kTile;
        tma_load_4d(k_s + s * kTileBytes, &tmap_k, kfull(s), 0, key0, kvh, b);
        tma_load_4d(k_s + s * kTileBytes + kHalfBytes, &tmap_k, kfull(s), 64, key0, kvh, b);
      }
      __syncwarp();
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    const int fmt = T::kDtype == KVC_BF16 ? 1 : 0;
    const uint32_t idesc1 = instr_desc(fmt, 128, 128), idesc2 = instr_desc(fmt, 128, nq);
    mbar_wait(qbar, 0);
    tc_fence_after();
    const uint64_t qd0 = smem_desc(q_s);
    int s = 0, ph = 0, buf = 0, bph = 1;
    for (int it = 0; it < n1 + n2; ++it) {
      const bool second = it >= n1;
      const uint32_t idesc = second ? idesc2 : idesc1;
      mbar_wait(tempty(buf), bph);
      mbar_wait(kfull(s), ph);
      tc_fence_after();
      const uint64_t kd0 = smem_desc(k_s + s * kTileBytes);
      const uint64_t ad0 = second ? kd0 : qd0, bd0 = second ? qd0 : kd0;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t off = (uint64_t)(((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4);
          tc_mma(tmem_base + buf * 128, ad0 + off, bd0 + off, idesc, k > 0);
        }
        tc_commit(tfull(buf));
        tc_commit(kempty(s));
      }
      __syncwarp();
      if (++buf == kHNumBuf) { buf = 0; bph ^= 1; }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    const int et = tid - 64;
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r_loc = quad * 32 + lane;
    const float scale = p.scale;
    unsigned my_gen = 0;
    if (et == 0) asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(my_gen) : "l"(&g_hbar_gen[slot]) : "memory");
    // stage the Q block: canonical K-major SWIZZLE_128B layout
#pragma unroll
    for (int it = 0; it < 128 * 16 / kHEpiThreads; ++it) {
      const int e = et + it * kHEpiThreads;
      const int r = e >> 4, c16 = e & 15;
      const int hl = r / W, i = r % W;
      const bool ok = hl < gl;
      const uint16_t* src = p.q + b * p.q_sb + (int64_t)(kvh * p.groups + head0 + (ok ? hl : 0)) * p.q_sh +
                            (int64_t)(nscore + i) * p.q_sl + c16 * 8;
      const uint32_t dst = q_s + (c16 >> 3) * kHalfBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(qbar);

    // Loop invariants the compiler would otherwise re-derive from the thread index and the shared-memory window on EVERY
    // tile (~35 of ~280 instructions per tile in the SASS): an empty asm makes each an opaque register value.
    uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + wg * kCols;
    uint32_t ebar = bar_s + 8u * (2 * S);            // tfull(i) = ebar + 8 i, tempty(i) = ebar + 8 (kHNumBuf + i)
    asm volatile("" : "+r"(tcol), "+r"(ebar));
    auto tfull = [&](int i) { return ebar + 8u * i; };
    auto tempty = [&](int i) { return ebar + 8u * (kHNumBuf + i); };
    const bool dead_rows = __all_sync(0xffffffffu, r_loc / W >= gl);
    // ---------------- pass 1 ----------------
    float m_run = -1e30f;
    float2 zacc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    int j = 0;
    for (; j < n1; ++j) {
      const int buf = j % kHNumBuf;
      const int key0 = tile_of(j) * kTile + wg * kCols;
      mbar_wait(tfull(buf), (j / kHNumBuf) & 1);
      tc_fence_after();
      if (dead_rows) {
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
        continue;
      }
      uint32_t va[32], vb[kChunks == 2 ? 32 : 1];
      tmem_ld32(tcol + buf * 128, va);
      if constexpr (kChunks == 2) tmem_ld32(tcol + buf * 128 + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));
      if (key0 + kCols > nscore) {
        const int last_key = min(p.L - 1, r_loc / W < gl ? nscore + r_loc % W : p.L - 1);
        p1_chunk<T, true>(va, scale, m_run, zacc, key0, last_key);
        if constexpr (kChunks == 2) p1_chunk<T, true>(vb, scale, m_run, zacc, key0 + 32, last_key);
      } else {
        p1_chunk<T, false>(va, scale, m_run, zacc, 0, 0);
        if constexpr (kChunks == 2) p1_chunk<T, false>(vb, scale, m_run, zacc, 0, 0);
      }
    }
    // ---------------- publish, grid barrier, fold ----------------
    {
      const float2 z2 = __fadd2_rn(__fadd2_rn(zacc[0], zacc[1]), __fadd2_rn(zacc[2], zacc[3]));
      xstat[wg * 128 + r_loc] = make_float2(m_run, z2.x + z2.y);
    }
    hepi_sync();
    if (et < 128) {
      float mn = xstat[et].x;
#pragma unroll
      for (int g = 1; g < NWG; ++g) mn = fmaxf(mn, xstat[g * 128 + et].x);
      float z = 0.f;
#pragma unroll
      for (int g = 0; g < NWG; ++g) { const float2 a = xstat[g * 128 + et]; z += a.y * fast_exp2((a.x - mn) * kLog2e); }
      p.stats[((int64_t)unit * p.nchunks + chunk) * 128 + et] = make_float2(mn, z);
    }
    __threadfence();
    hepi_sync();
    if (et == 0) {
      // sense-reversing barrier on the library-owned words of this launch's slot (one slot per stream:
      // launches of one stream are serialised, launches of different streams use different slots)
      const unsigned total = gridDim.x * gridDim.y;
      const unsigned old = atomicAdd(&g_hbar_count[slot], 1u);
      if (old == total - 1) {
        g_hbar_count[slot] = 0u;
        __threadfence();
        atomicAdd(&g_hbar_gen[slot], 1u);
      } else {
        unsigned gen = my_gen;
        for (unsigned spin = 0; gen == my_gen; ++spin) {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(&g_hbar_gen[slot]) : "memory");
          if (gen == my_gen) {
            __nanosleep(32);
            if (spin > (1u << 24)) {
              printf("kvc scores_fused(half): grid barrier timed out (block %d,%d slot %d)\n", blockIdx.x, blockIdx.y, slot);
              __trap();
            }
          }
        }
      }
    }
    hepi_sync();
    {
      constexpr int kParts = kHEpiThreads / 128;
      const int r = et % 128, q = et / 128;
      const float2* st = p.stats + (int64_t)unit * p.nchunks * 128 + r;
      float M = -1e30f, Z = 0.f;
      for (int c0 = q; c0 < p.nchunks; c0 += 8 * kParts) {
        float2 sv[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          sv[u] = make_float2(-1e30f, 0.f);
          if (c0 + u * kParts < p.nchunks) {
            const float2* a = st + (int64_t)(c0 + u * kParts) * 128;
            asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(sv[u].x), "=f"(sv[u].y) : "l"(a) : "memory");
          }
        }
        float mx = M;
#pragma unroll
        for (int u = 0; u < 8; ++u) mx = fmaxf(mx, sv[u].x);
        Z *= fast_exp2((M - mx) * kLog2e);
#pragma unroll
        for (int u = 0; u < 8; ++u) Z += sv[u].y * fast_exp2((sv[u].x - mx) * kLog2e);
        M = mx;
      }
      xstat[q * 128 + r] = make_float2(M, Z);
      hepi_sync();
      if (et < 128) {
        float mx = xstat[et].x;
#pragma unroll
        for (int g = 1; g < kParts; ++g) mx = fmaxf(mx, xstat[g * 128 + et].x);
        float z = 0.f;
#pragma unroll
        for (int g = 0; g < kParts; ++g) { const float2 a = xstat[g * 128 + et]; z += a.y * fast_exp2((a.x - mx) * kLog2e); }
        crow[et] = (et / W < gl) ? (-mx * kLog2e - log2f(z)) : -INFINITY;
      }
    }
    hepi_sync();
    // ---------------- pass 2 ----------------
    constexpr int kHeadsPerChunk = 32 / W;
    // scores row of the first head of this warpgroup's first chunk (the heads of a chunk are consecutive rows)
    uint16_t* out_row = p.out + ((int64_t)b * p.Hq + kvh * p.groups + head0 + (wg * kCols) / W) * (int64_t)nscore;
    asm volatile("" : "+l"(out_row));
    for (; j < n1 + n2; ++j) {
      const int buf = j % kHNumBuf;
      const int key = tile_of(j) * kTile + r_loc;
      mbar_wait(tfull(buf), (j / kHNumBuf) & 1);
      tc_fence_after();
      if ((wg * kCols) / W >= gl) {  // all chunks of this warpgroup lie beyond the last head
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));
        continue;
      }
      uint32_t va[32], vb[kChunks == 2 ? 32 : 1];
      tmem_ld32(tcol + buf * 128, va);
      if constexpr (kChunks == 2) tmem_ld32(tcol + buf * 128 + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty(buf));
#pragma unroll
      for (int c = 0; c < kChunks; ++c) {
        const int col0 = wg * kCols + c * 32;
        if (col0 / W >= gl) continue;
        float2 acc[kHeadsPerChunk];
#pragma unroll
        for (int h = 0; h < kHeadsPerChunk; ++h) acc[h] = make_float2(0.f, 0.f);
        if constexpr (kChunks == 2) p2_chunk<T, W>(c == 0 ? va : vb, scale, crow + col0, acc);
        else p2_chunk<T, W>(va, scale, crow + col0, acc);
        if (key < nscore) {
#pragma unroll
          for (int h = 0; h < kHeadsPerChunk; ++h) {
            const int head = col0 / W + h;
            if (head < gl) {
              float sum = acc[h].x + acc[h].y;
              if (p.reduce_mean) sum = div_exact(sum, (float)p.w, 1.0f / (float)p.w);
              out_row[(int64_t)(c * kHeadsPerChunk + h) * nscore + key] = T::round1(sum);
            }
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  tl_end(p.tl);
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(256u) : "memory");
  }
}

// Barrier slot of a stream: launches on one stream never overlap, launches on different streams must not
// share barrier words.  More than kHSlots distinct streams are refused.
static int half_slot(cudaStream_t stream) {
  static std::mutex mu;
  static cudaStream_t seen[64][kHSlots];   // the barrier words are device globals: one table per device
  static int n[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n[dev]; ++i)
    if (seen[dev][i] == stream) return i;
  if (n[dev] == kHSlots) return -1;
  seen[dev][n[dev]] = stream;
  return n[dev]++;
}

template <typename T, int W, int NWG>
static int launch_half(const CUtensorMap& tmap, const Params& sp, int units, bool coop, cudaStream_t stream) {
  auto kern = window_score_half_kernel<T, W, NWG>;
  const int smem = HalfSmem::kAlloc;
  KVC_SET_MAX_SMEM(kern, smem);
  const int slot = half_slot(stream);
  KVC_REQUIRE(slot >= 0, KVC_ERR_UNSUPPORTED, "half-footprint tcgen05 scoring kernel: more than %d distinct streams on this device", kHSlots);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sp.nchunks, units);
  cfg.blockDim = dim3(64 + 128 * NWG);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  // no programmatic dependent launch here: with two launches in flight a pre-launched successor of the same
  // stream competes with the other stream's launch for the free CTA slot of each SM (measured: 1.49 vs 1.46 ms)
  // coop (KVC_IMPL_AUTO on shapes with several row blocks per KV head): cooperative launch, co-residency of the
  // grid guaranteed by the driver whatever else the caller has in flight
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  if (coop) { cfg.attrs = attr; cfg.numAttrs = 1; }
  KVC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap, sp, slot));
  return launch_status("window_score_half_kernel");
}

template <typename T>
static int dispatch_half(const CUtensorMap& tmap, const Params& sp, int units, bool coop, cudaStream_t stream) {
  // four warpgroups (56 registers, 8 epilogue warps per SM sub-partition with two CTAs on the SM) measured 29.0 vs
  // 29.9 us per launch with two launches in flight, 1.431 vs 1.452 ms per cfg-2 step; KVC_HALF_WG=2 selects the
  // two-warpgroup build
  static const bool four = [] { const char* e = getenv("KVC_HALF_WG"); return !(e && atoi(e) == 2); }();
  if (four) {
    switch (sp.w) {
      case 8: return launch_half<T, 8, 4>(tmap, sp, units, coop, stream);
      case 16: return launch_half<T, 16, 4>(tmap, sp, units, coop, stream);
      default: return launch_half<T, 32, 4>(tmap, sp, units, coop, stream);
    }
  }
  switch (sp.w) {
    case 8: return launch_half<T, 8, 2>(tmap, sp, units, coop, stream);
    case 16: return launch_half<T, 16, 2>(tmap, sp, units, coop, stream);
    default: return launch_half<T, 32, 2>(tmap, sp, units, coop, stream);
  }
}

struct Plan {
  int rb, chunk_tiles, nchunks;
  size_t stats_bytes;
};

static Plan make_plan(int B, int Hq, int Hkv, int L, int w, bool split_rb = false) {
  Plan pl{};
  const int g = Hq / Hkv;
  pl.rb = (g * w + 127) / 128;
  const int units = B * Hkv * (split_rb ? pl.rb : 1);   // the half-footprint kernel takes every row block as a unit
  const int tiles = (L + kTile - 1) / kTile;
  int cu = sm_count() / units;
  if (cu < 1) cu = 1;
  pl.chunk_tiles = (tiles + cu - 1) / cu;
  pl.nchunks = (tiles + pl.chunk_tiles - 1) / pl.chunk_tiles;
  pl.stats_bytes = ((size_t)units * pl.nchunks * (split_rb ? 1 : pl.rb) * 128 * sizeof(float2) + 255) & ~(size_t)255;
  return pl;
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

template <typename T, int RB, int W>
static int launch(const CUtensorMap& tmap, const Params& sp, int units, cudaStream_t stream) {
  auto kern = window_score_fused_kernel<T, RB, W>;
  const int smem = Smem<RB>::kAlloc;
  KVC_SET_MAX_SMEM(kern, smem);
  // cooperative launch: the grid is scheduled only when every CTA can be resident, so two grids
  // with barriers (other streams / processes) can never hold each other's SMs hostage
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(sp.nchunks, units);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  cfg.attrs = attr;
  // KVC_FUSED_COOP=0: plain launch.  CTAs then trickle onto SMs as other streams' kernels (e.g. the
  // asynchronous select/gather tail) retire instead of waiting for a fully free GPU; only safe when
  // no second barrier-synchronised grid can run concurrently on the device.
  static const bool coop = [] { const char* e = getenv("KVC_FUSED_COOP"); return !(e && e[0] == '0'); }();
  static const bool pdl = [] { const char* e = getenv("KVC_PDL"); return !(e && e[0] == '0'); }();
  // the cooperative attribute and programmatic serialization do not combine: PDL only on plain launches
  if (coop) { attr[0] = attr[1]; cfg.numAttrs = 1; } else { cfg.numAttrs = pdl ? 1 : 0; }
  KVC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap, sp));
  return launch_status("window_score_fused_kernel");
}

template <typename T, int RB>
static int dispatch_w(const CUtensorMap& tmap, const Params& sp, int units, cudaStream_t stream) {
  switch (sp.w) {
    case 8: return launch<T, RB, 8>(tmap, sp, units, stream);
    case 16: return launch<T, RB, 16>(tmap, sp, units, stream);
    default: return launch<T, RB, 32>(tmap, sp, units, stream);
  }
}

}  // namespace tcf

#ifdef KVC_TRACE
extern "C" int kvc_debug_ftrace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, tcf::g_trace, sizeof(long long) * (size_t)n);
}
extern "C" int kvc_debug_tile_trace(unsigned* out) {
  return (int)cudaMemcpyFromSymbol(out, tcf::g_tile_trace, sizeof(unsigned) * 4 * 64 * 8);
}
#endif

bool window_scores_tc_supported(int B, int Hq, int Hkv, int L, int d, int w);

// The grid barrier needs every CTA resident at once: units * chunks must fit the SM count.
bool window_scores_fused_supported(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_tc_supported(B, Hq, Hkv, L, d, w)) return false;
  if (w > 32) return false;  // a 64-row head spans two warpgroups' columns: the two-launch kernel covers it
  const tcf::Plan pl = tcf::make_plan(B, Hq, Hkv, L, w);
  return B * Hkv * pl.nchunks <= sm_count();
}

// Shapes with several 128-row blocks per KV head (70B: 8 query heads x 32 window rows): the half-footprint kernel with
// one unit per row block is faster than the two-accumulator full-size kernel (37.7 vs 41.0 us for one KV head at
// 128K) -- KVC_IMPL_AUTO takes it when its grid fits.
bool window_scores_split_preferred(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_fused_supported(B, Hq, Hkv, L, d, w) || 128 % w) return false;
  const tcf::Plan pl = tcf::make_plan(B, Hq, Hkv, L, w, true);
  return pl.rb >= 2 && B * Hkv * pl.rb * pl.nchunks <= sm_count();
}

size_t window_scores_fused_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_tc_supported(B, Hq, Hkv, L, d, w)) return 0;
  const size_t a = tcf::make_plan(B, Hq, Hkv, L, w).stats_bytes, b = tcf::make_plan(B, Hq, Hkv, L, w, true).stats_bytes;
  return (a > b ? a : b) + 256;
}

int window_scores_fused(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                        int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce,
                        void* out_scores, void* workspace, size_t workspace_bytes, int half, cudaStream_t stream) {
  using namespace tcf;
  KVC_REQUIRE(window_scores_fused_supported(B, Hq, Hkv, L, d, w), KVC_ERR_UNSUPPORTED,
              "fused tcgen05 scoring kernel does not cover this shape (Hq/Hkv=%d, w=%d, d=%d, units=%d)", Hq / Hkv, w, d,
              B * Hkv);
  KVC_REQUIRE(((q_sb | q_sh | q_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel: query rows must be 16-byte aligned");
  KVC_REQUIRE(((k_sb | k_sh | k_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel: key rows must be 16-byte aligned (TMA)");
  // half: 0 = full-size kernel, 1 = half-footprint kernel (plain launch: the caller keeps at most two in flight),
  // 2 = half-footprint kernel launched cooperatively
  const Plan pl = make_plan(B, Hq, Hkv, L, w, half != 0);
  KVC_REQUIRE(workspace_bytes >= pl.stats_bytes + 256, KVC_ERR_WORKSPACE, "window_scores workspace: have %zu need %zu",
              workspace_bytes, pl.stats_bytes + 256);
  EncodeTiledFn enc = encode_fn();
  KVC_REQUIRE(enc != nullptr, KVC_ERR_DEVICE, "cuTensorMapEncodeTiled is not available from this driver");
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)d, (cuuint64_t)L, (cuuint64_t)Hkv, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)k_sl * 2, (cuuint64_t)k_sh * 2, (cuuint64_t)k_sb * 2};
  if (Hkv == 1) gstr[1] = gstr[0] * (cuuint64_t)L;
  if (B == 1) gstr[2] = gstr[0] * (cuuint64_t)L * (cuuint64_t)Hkv;
  const cuuint32_t box[4] = {64, (cuuint32_t)kTile, 1, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult cr = enc(&tmap, dtype == KVC_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16,
                          4, const_cast<void*>(k), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  KVC_REQUIRE(cr == CUDA_SUCCESS, KVC_ERR_UNSUPPORTED, "cuTensorMapEncodeTiled failed with CUresult %d", (int)cr);
  Params sp{};
  sp.q = (const uint16_t*)q; sp.q_sb = q_sb; sp.q_sh = q_sh; sp.q_sl = q_sl;
  sp.B = B; sp.Hq = Hq; sp.Hkv = Hkv; sp.L = L; sp.w = w; sp.groups = Hq / Hkv;
  sp.chunk_tiles = pl.chunk_tiles; sp.nchunks = pl.nchunks;
  sp.scale = 1.0f / sqrtf((float)d);
  sp.reduce_mean = reduce == KVC_REDUCE_MEAN;
  sp.stats = (float2*)workspace;
  sp.out = (uint16_t*)out_scores;
  sp.tl = KVC_TL_SLOT(0);
  const int units = B * Hkv;
  if (half) {
    KVC_REQUIRE(128 % w == 0, KVC_ERR_UNSUPPORTED, "half-footprint tcgen05 scoring kernel: window %d does not divide 128", w);
    KVC_REQUIRE(units * pl.rb * pl.nchunks <= sm_count(), KVC_ERR_UNSUPPORTED,
                "half-footprint tcgen05 scoring kernel: %d CTAs exceed the %d SMs (two launches must be co-resident)",
                units * pl.rb * pl.nchunks, sm_count());
    sp.rbs = pl.rb;
    return dtype == KVC_BF16 ? dispatch_half<Bf16>(tmap, sp, units * pl.rb, half == 2, stream)
                             : dispatch_half<Fp16>(tmap, sp, units * pl.rb, half == 2, stream);
  }
  if (dtype == KVC_BF16)
    return pl.rb == 1 ? dispatch_w<Bf16, 1>(tmap, sp, units, stream) : dispatch_w<Bf16, 2>(tmap, sp, units, stream);
  return pl.rb == 1 ? dispatch_w<Fp16, 1>(tmap, sp, units, stream) : dispatch_w<Fp16, 2>(tmap, sp, units, stream);
}

}  // namespace kvc
