// Window scoring on Blackwell tensor cores: TMA-staged K tiles, tcgen05.mma with TMEM
// accumulators, warp-specialised (1 TMA warp, 1 MMA warp, 8 epilogue warps).
//
// Work decomposition is the same as scores_mma.cu: a CTA owns one chunk of keys of one
// (batch, KV head) unit and stacks the unit's g = Hq/Hkv query heads into 128-row blocks
// (row = head_in_block * w + window_row), so K is streamed once per group.
//
//   pass 1 (stats)   D[128 query rows x 128 keys] = Q_blk . Ktile^T      thread <-> query row
//                    per row: running max m and Z = sum_j exp(l_ij - m) over the chunk
//   pass 2 (scores)  D[128 keys x NQ query rows]  = Ktile . Q_blk^T      thread <-> key
//                    p_ij = exp(l_ij - M_i)/Z_i -> model dtype, summed over each head's w rows
//
// Swapping the operand roles between the passes makes both reductions thread-local: the row
// statistics never cross lanes in pass 1 and the per-key window sums never cross lanes in pass 2.
// Both operands are K-major 128-byte-swizzled shared-memory tiles, so the same TMA-written K
// stage (two [128 x 64] halves) serves as B in pass 1 and as A in pass 2.
//
// Rounding points of the reference (pyramidkv_utils.py:317-327) are reproduced exactly as in
// scores_mma.cu: GEMM result -> dtype, * 1/sqrt(d) in fp32 -> dtype, softmax in fp32 -> dtype,
// window sum in fp32 -> dtype.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace kvc {
namespace tc {

// ---- timing trace (KVC_TC_DEBUG & 4): clock64 stamps of CTA (0,0), read back by kvc_debug_trace ----
__device__ long long g_trace[512];
#define KVC_TRACE(slot) do { if ((p.dbg & 4) && (((p.dbg & 8) != 0) != kPass2) && blockIdx.x == 0 && blockIdx.y == 0) g_trace[(slot)] = clock64(); } while (0)

constexpr int kTile = 128;                    // keys per tile
constexpr int kD = 128;                       // head_dim
constexpr int kHalfBytes = kTile * 64 * 2;    // one [128 x 64] swizzled half: 16 KiB
constexpr int kTileBytes = 2 * kHalfBytes;    // 32 KiB
constexpr int kNumBuf = 4;                    // TMEM accumulator buffers (128 columns each)
constexpr int kEpiWarps = 8;
constexpr int kNumWG = kEpiWarps / 4;
constexpr int kThreads = 32 * (2 + kEpiWarps);

struct Params {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  int B, Hq, Hkv, L, w, groups;
  int chunk_tiles, nchunks;
  float scale;
  int reduce_mean;
  int dbg;         // KVC_TC_DEBUG bits: 4 = record the clock64 trace of CTA (0,0) for pass 2, 4|8 = for pass 1
  float2* stats;   // [units][nchunks][RB*128]
  uint16_t* out;   // [B,Hq,L-w]
};

// ---- PTX wrappers ---------------------------------------------------------------------------
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
// Bounded wait: a protocol bug must trap, not hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  for (uint32_t spin = 0; !done; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (!done && spin > (1u << 22)) {
      printf("kvc scores_tc: mbarrier wait timed out (block %d,%d thread %d bar %u parity %u)\n", blockIdx.x,
             blockIdx.y, threadIdx.x, bar, parity);
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
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
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

// K-major, SWIZZLE_128B shared-memory operand descriptor (cute::UMMA::SmemDescriptor):
// start address >> 4 in [0,14), LBO (ignored for swizzled K-major) = 1 in [16,30), SBO = 1024 B
// (8 rows x 128 B) >> 4 in [32,46), version = 1 in [46,48), layout SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
// kind::f16 instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (1 << 4), A/B format at
// [7,10)/[10,13) (F16 = 0, BF16 = 1), both K-major, N >> 3 at [17,23), M >> 4 at [24,29).
__device__ __forceinline__ uint32_t instr_desc(int fmt, int M, int N) {
  return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

template <int RB>
struct Smem {
  static constexpr int kStages = RB == 1 ? 5 : 4;
  static constexpr int kQBytes = RB * kTileBytes;
  static constexpr int kKBytes = kStages * kTileBytes;
  static constexpr int kBarOff = kQBytes + kKBytes;                    // 8-byte barriers
  static constexpr int kNumBars = 2 * kStages + 2 * kNumBuf + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kCrowOff = (kTmemPtrOff + 4 + 15) & ~15;        // float[RB*128]
  static constexpr int kXstatOff = kCrowOff + RB * 128 * 4;            // float2[RB][2][128]
  static constexpr int kTotal = kXstatOff + RB * kNumWG * 128 * 8;
  static constexpr int kAlloc = kTotal + 1024;                          // slack for the 1024-B alignment
};

template <typename T, bool kPass2, int RB, int W>
__global__ void __launch_bounds__(kThreads, 1)
window_score_tc_kernel(const __grid_constant__ CUtensorMap tmap_k, const Params p) {
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
  const int key_limit = kPass2 ? nscore : p.L;
  const int tiles_total = (key_limit + kTile - 1) / kTile;
  const int tile_begin = chunk * p.chunk_tiles;
  const int ntiles = min(tile_begin + p.chunk_tiles, tiles_total) - tile_begin;
  if (ntiles <= 0) return;
  if (threadIdx.x == 0) KVC_TRACE(0);
  const int nq = RB == 1 ? p.groups * p.w : 128;  // query rows per row block that carry data (UMMA N in pass 2)

  const uint32_t q_s = base, k_s = base + SM::kQBytes, bar_s = base + SM::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (S + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * S + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * S + kNumBuf + i); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::kTmemPtrOff);
  float* crow = reinterpret_cast<float*>(smem + SM::kCrowOff);
  float2* xstat = reinterpret_cast<float2*>(smem + SM::kXstatOff);

  // ---- one-time setup: barriers + TMEM, then the roles start; the K stream does not wait for Q ----
  const uint32_t qbar = bar_s + 8u * (2 * S + 2 * kNumBuf);
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(kfull(s), 1); mbar_init(kempty(s), 1); }
    for (int i = 0; i < kNumBuf; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
    mbar_init(qbar, kEpiWarps * 32);
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
  if (threadIdx.x == 0) KVC_TRACE(1);

  if (warp >= 2) {
    // Epilogue warps stage the Q block(s) with cp.async (all 16-byte chunks in flight at once) in the
    // canonical K-major SWIZZLE_128B layout:
    //   byte(row, col) = half(col/64) * 16 KiB + row * 128 + ((col%64)/8 ^ row%8) * 16 + (col%8) * 2
    const int et = tid - 64;
#pragma unroll
    for (int it = 0; it < (RB * 128 * 16 + kEpiWarps * 32 - 1) / (kEpiWarps * 32); ++it) {
      const int e = et + it * (kEpiWarps * 32);
      if (e >= RB * 128 * 16) break;
      const int rg = e >> 4, c16 = e & 15;  // stacked query row, 16-byte chunk of the row
      const int rb = rg >> 7, r = rg & 127;
      const int hl = rg / W, i = rg % W;
      const bool ok = hl < p.groups;
      const uint16_t* src = p.q + b * p.q_sb + (int64_t)(kvh * p.groups + (ok ? hl : 0)) * p.q_sh +
                            (int64_t)(nscore + i) * p.q_sl + c16 * 8;
      const uint32_t dst = q_s + rb * kTileBytes + (c16 >> 3) * kHalfBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
      const int sz = ok ? 16 : 0;  // padding rows are zero-filled
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if constexpr (kPass2) {
      // fold every chunk's (m, Z) of this unit into c_row = -M*log2e - log2 Z (loads batched by 8)
      for (int r = et; r < RB * 128; r += kEpiWarps * 32) {
        const float2* st = p.stats + (int64_t)unit * p.nchunks * (RB * 128) + r;
        float M = -1e30f, Z = 0.f;
        for (int c0 = 0; c0 < p.nchunks; c0 += 8) {
          float2 sv[8];
#pragma unroll
          for (int u = 0; u < 8; ++u)
            sv[u] = c0 + u < p.nchunks ? st[(int64_t)(c0 + u) * (RB * 128)] : make_float2(-1e30f, 0.f);
          float mx = M;
#pragma unroll
          for (int u = 0; u < 8; ++u) mx = fmaxf(mx, sv[u].x);
          Z *= fast_exp2((M - mx) * kLog2e);
#pragma unroll
          for (int u = 0; u < 8; ++u) Z += sv[u].y * fast_exp2((sv[u].x - mx) * kLog2e);
          M = mx;
        }
        crow[r] = (r / W < p.groups) ? (-M * kLog2e - log2f(Z)) : -INFINITY;
      }
    }
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy Q writes -> visible to UMMA
    mbar_arrive(qbar);
    asm volatile("bar.sync 1, %0;" ::"n"(kEpiWarps * 32) : "memory");  // c_row visible to all epilogue warps
  }

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        mbar_wait(kempty(s), ((t / S) & 1) ^ 1);
        mbar_expect_tx(kfull(s), kTileBytes);
        const int key0 = (tile_begin + t) * kTile;
        tma_load_4d(k_s + s * kTileBytes, &tmap_k, kfull(s), 0, key0, kvh, b);
        tma_load_4d(k_s + s * kTileBytes + kHalfBytes, &tmap_k, kfull(s), 64, key0, kvh, b);
        if (t < 32) KVC_TRACE(32 + t);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc(T::kDtype == KVC_BF16 ? 1 : 0, 128, kPass2 ? nq : 128);
      mbar_wait(qbar, 0);  // Q block(s) staged
      KVC_TRACE(2);
      tc_fence_after();
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        mbar_wait(kfull(s), (t / S) & 1);
        if (t < 32) KVC_TRACE(64 + t);
        tc_fence_after();
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          const int j = t * RB + rb, buf = j % kNumBuf;
          mbar_wait(tempty(buf), ((j / kNumBuf) & 1) ^ 1);
          tc_fence_after();
          const uint32_t ks = k_s + s * kTileBytes, qs = q_s + rb * kTileBytes;
#pragma unroll
          for (int k = 0; k < kD / 16; ++k) {
            const uint32_t off = (k >> 2) * kHalfBytes + (k & 3) * 32;  // 16 elements = 32 B inside the swizzle atom
            const uint64_t kd = smem_desc(ks + off), qd = smem_desc(qs + off);
            tc_mma(tmem_base + buf * 128, kPass2 ? kd : qd, kPass2 ? qd : kd, idesc, k > 0);
          }
          tc_commit(tfull(buf));
        }
        tc_commit(kempty(s));
      }
    }
  } else {
    // ===== epilogue warps: kNumWG groups of four warps take accumulator tiles round-robin =====
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int r_loc = quad * 32 + lane;        // pass 1: query row in block; pass 2: key in tile
    float m_run[RB], z_run[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) { m_run[rb] = -1e30f; z_run[rb] = 0.f; }

    for (int j = wg; j < ntiles * RB; j += kNumWG) {
      const int t = j / RB, rb = j % RB, buf = j % kNumBuf;
      const int key0 = (tile_begin + t) * kTile;
      mbar_wait(tfull(buf), (j / kNumBuf) & 1);
      if (warp == 2 + 4 * wg && lane == 0 && j < 32) KVC_TRACE(96 + j);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128;

      if constexpr (!kPass2) {
        // ---- stats: this thread owns one query row, the 128 columns are keys ----
        const int rg = rb * 128 + r_loc;
        const int hl = rg / W, wi = rg % W;
        const bool needs_mask = key0 + kTile > nscore;
        const int last_key = min(p.L - 1, hl < p.groups ? nscore + wi : p.L - 1);  // last visible key of this row
        // packed fp32x2 math (FMUL2 / FFMA2 / FADD2) halves the issue slots of the scale/exp chain
        const float2 scale2 = make_float2(p.scale, p.scale);
        uint32_t r2[64];
        {
          // all four 32-column loads in flight, ONE wait, then the accumulator buffer goes straight back
          uint32_t v[4][32];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld32(taddr + cc * 32, v[cc]);
          tmem_ld_wait();
          if (warp == 2 + 4 * wg && lane == 0 && j < 32) KVC_TRACE(160 + j);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
          // staged so that the 16 independent pack -> scale -> pack chains of a chunk interleave
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t p1[16];
            float2 f[16];
#pragma unroll
            for (int pi = 0; pi < 16; ++pi)
              p1[pi] = T::pack(__uint_as_float(v[cc][2 * pi]), __uint_as_float(v[cc][2 * pi + 1]));
#pragma unroll
            for (int pi = 0; pi < 16; ++pi) f[pi] = __fmul2_rn(make_float2(T::lo(p1[pi]), T::hi(p1[pi])), scale2);
#pragma unroll
            for (int pi = 0; pi < 16; ++pi) r2[cc * 16 + pi] = T::pack(f[pi].x, f[pi].y);
          }
        }
        if (needs_mask) {  // only the tile(s) touching the window block / the ragged end
#pragma unroll
          for (int i = 0; i < 64; ++i) {
            const int key = key0 + 2 * i;
            if (key > last_key) r2[i] = (r2[i] & 0xffff0000u) | (T::kNegInfPair & 0xffffu);
            if (key + 1 > last_key) r2[i] = (r2[i] & 0x0000ffffu) | (T::kNegInfPair & 0xffff0000u);
          }
        }

        if (warp == 2 + 4 * wg && lane == 0 && j < 32) KVC_TRACE(192 + j);
        uint32_t pm = r2[0];
#pragma unroll
        for (int i = 1; i < 64; ++i) pm = T::max2(pm, r2[i]);
        const float m_new = fmaxf(m_run[rb], fmaxf(T::lo(pm), T::hi(pm)));
        const float2 ml2 = make_float2(-m_new * kLog2e, -m_new * kLog2e);
        const float2 l2e2 = make_float2(kLog2e, kLog2e);
        float2 zacc[4];
        zacc[0] = make_float2(z_run[rb] * fast_exp2((m_run[rb] - m_new) * kLog2e), 0.f);
        zacc[1] = zacc[2] = zacc[3] = make_float2(0.f, 0.f);
        if (warp == 2 + 4 * wg && lane == 0 && j < 32) KVC_TRACE(224 + j);
#pragma unroll
        for (int bt = 0; bt < 4; ++bt) {  // 16 pairs per batch: FFMA2s, then the 32 exps, then the adds
          float2 a[16];
#pragma unroll
          for (int i = 0; i < 16; ++i)
            a[i] = __ffma2_rn(make_float2(T::lo(r2[bt * 16 + i]), T::hi(r2[bt * 16 + i])), l2e2, ml2);
#pragma unroll
          for (int i = 0; i < 16; ++i) a[i] = make_float2(fast_exp2(a[i].x), fast_exp2(a[i].y));
#pragma unroll
          for (int i = 0; i < 16; ++i) zacc[i & 3] = __fadd2_rn(zacc[i & 3], a[i]);
        }
        const float2 z2 = __fadd2_rn(__fadd2_rn(zacc[0], zacc[1]), __fadd2_rn(zacc[2], zacc[3]));
        m_run[rb] = m_new;
        if (warp == 2 + 4 * wg && lane == 0 && j < 32) KVC_TRACE(128 + j);
        z_run[rb] = z2.x + z2.y;
      } else {
        // ---- scores: this thread owns one key, the columns are the stacked query rows ----
        constexpr int kHeadsPerBlk = 128 / W;
        float2 acc[kHeadsPerBlk];
#pragma unroll
        for (int h = 0; h < kHeadsPerBlk; ++h) acc[h] = make_float2(0.f, 0.f);
        const float2 scale2 = make_float2(p.scale, p.scale);
        const float2 l2e2 = make_float2(kLog2e, kLog2e);
        const float* cr = crow + rb * 128;
        // All four 32-column chunks are processed: columns >= nq belong to padding rows whose
        // c_row is -inf, so they contribute exp2(-inf) = 0 (and their heads are never stored).
        uint32_t r2[64];
        {
          uint32_t v[4][32];
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) tmem_ld32(taddr + cc * 32, v[cc]);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));  // accumulator drained: release it before the exp work
#pragma unroll
          for (int cc = 0; cc < 4; ++cc) {
            uint32_t p1[16];
            float2 f[16];
#pragma unroll
            for (int pi = 0; pi < 16; ++pi)
              p1[pi] = T::pack(__uint_as_float(v[cc][2 * pi]), __uint_as_float(v[cc][2 * pi + 1]));
#pragma unroll
            for (int pi = 0; pi < 16; ++pi) f[pi] = __fmul2_rn(make_float2(T::lo(p1[pi]), T::hi(p1[pi])), scale2);
#pragma unroll
            for (int pi = 0; pi < 16; ++pi) r2[cc * 16 + pi] = T::pack(f[pi].x, f[pi].y);
          }
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {  // 16 pairs per batch: FFMA2s, 32 exps, packs, adds
          float2 a[16];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cr + cc * 32 + q4 * 4);
            a[2 * q4] = __ffma2_rn(make_float2(T::lo(r2[cc * 16 + 2 * q4]), T::hi(r2[cc * 16 + 2 * q4])), l2e2,
                                   make_float2(c4.x, c4.y));
            a[2 * q4 + 1] = __ffma2_rn(make_float2(T::lo(r2[cc * 16 + 2 * q4 + 1]), T::hi(r2[cc * 16 + 2 * q4 + 1])),
                                       l2e2, make_float2(c4.z, c4.w));
          }
          uint32_t pp[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) pp[i] = T::pack(fast_exp2(a[i].x), fast_exp2(a[i].y));
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const int col = cc * 32 + 2 * i;
            acc[col / W] = __fadd2_rn(acc[col / W], make_float2(T::lo(pp[i]), T::hi(pp[i])));
          }
        }
        const int key = key0 + r_loc;
        if (key < nscore) {
#pragma unroll
          for (int h = 0; h < kHeadsPerBlk; ++h) {
            const int head = rb * kHeadsPerBlk + h;
            if (head < p.groups) {
              float s = acc[h].x + acc[h].y;
              if (p.reduce_mean) s = div_exact(s, (float)p.w, 1.0f / (float)p.w);
              p.out[((int64_t)b * p.Hq + kvh * p.groups + head) * nscore + key] = T::round1(s);
            }
          }
        }
      }
    }
    if constexpr (!kPass2) {
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) xstat[(rb * kNumWG + wg) * 128 + r_loc] = make_float2(m_run[rb], z_run[rb]);
    }
  }

  if (threadIdx.x == 0) KVC_TRACE(3);
  tc_fence_before();
  __syncthreads();
  if constexpr (!kPass2) {
    for (int r = tid; r < RB * 128; r += kThreads) {
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
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ---- host side ---------------------------------------------------------------------------------
struct Plan {
  int rb, chunk_tiles, nchunks;
  size_t stats_bytes;
};

static Plan make_plan(int B, int Hq, int Hkv, int L, int w) {
  Plan pl{};
  const int g = Hq / Hkv;
  pl.rb = (g * w + 127) / 128;
  const int units = B * Hkv;
  const int tiles = (L + kTile - 1) / kTile;
  int cu = sm_count() / units;  // CTAs per unit: one CTA per SM (shared memory bound), single wave
  if (cu < 1) cu = 1;
  pl.chunk_tiles = (tiles + cu - 1) / cu;
  pl.nchunks = (tiles + pl.chunk_tiles - 1) / pl.chunk_tiles;
  pl.stats_bytes = (size_t)units * pl.nchunks * pl.rb * 128 * sizeof(float2);
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
static int launch_pair(const CUtensorMap& tmap, const Params& sp, int units, cudaStream_t stream) {
  auto k1 = window_score_tc_kernel<T, false, RB, W>;
  auto k2 = window_score_tc_kernel<T, true, RB, W>;
  const int smem = Smem<RB>::kAlloc;
  KVC_SET_MAX_SMEM(k1, smem);
  KVC_SET_MAX_SMEM(k2, smem);
  dim3 grid(sp.nchunks, units);
  k1<<<grid, kThreads, smem, stream>>>(tmap, sp);
  if (int rc = launch_status("window_score_tc_kernel<stats>")) return rc;
  k2<<<grid, kThreads, smem, stream>>>(tmap, sp);
  return launch_status("window_score_tc_kernel<scores>");
}

template <typename T, int RB>
static int dispatch_w(const CUtensorMap& tmap, const Params& sp, int units, cudaStream_t stream) {
  switch (sp.w) {
    case 8: return launch_pair<T, RB, 8>(tmap, sp, units, stream);
    case 16: return launch_pair<T, RB, 16>(tmap, sp, units, stream);
    case 32: return launch_pair<T, RB, 32>(tmap, sp, units, stream);
    default: return launch_pair<T, RB, 64>(tmap, sp, units, stream);
  }
}

}  // namespace tc

bool window_scores_tc_supported(int B, int Hq, int Hkv, int L, int d, int w) {
  (void)B; (void)L;
  if (d != 128 || Hkv <= 0 || Hq % Hkv) return false;
  if (!(w == 8 || w == 16 || w == 32 || w == 64)) return false;
  const int rows = (Hq / Hkv) * w;
  return rows % 32 == 0 && rows <= 256;
}

size_t window_scores_tc_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_tc_supported(B, Hq, Hkv, L, d, w)) return 0;
  return (tc::make_plan(B, Hq, Hkv, L, w).stats_bytes + 255) & ~(size_t)255;
}

int window_scores_tc(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb, int64_t k_sh,
                     int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce, void* out_scores,
                     void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  using namespace tc;
  KVC_REQUIRE(window_scores_tc_supported(B, Hq, Hkv, L, d, w), KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel does not cover Hq/Hkv=%d, w=%d, d=%d", Hq / Hkv, w, d);
  KVC_REQUIRE(((q_sb | q_sh | q_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel: query rows must be 16-byte aligned");
  KVC_REQUIRE(((k_sb | k_sh | k_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel: key rows must be 16-byte aligned (TMA)");
  const Plan pl = make_plan(B, Hq, Hkv, L, w);
  KVC_REQUIRE(workspace_bytes >= pl.stats_bytes, KVC_ERR_WORKSPACE, "window_scores workspace: have %zu need %zu",
              workspace_bytes, pl.stats_bytes);
  EncodeTiledFn enc = encode_fn();
  KVC_REQUIRE(enc != nullptr, KVC_ERR_DEVICE, "cuTensorMapEncodeTiled is not available from this driver");

  // K as a 4-D tensor {d, L, Hkv, B}; box = one 64-column half of a 128-key tile, 128-byte swizzle,
  // out-of-range keys (ragged last tile) are zero-filled by the TMA unit.
  CUtensorMap tmap;
  const cuuint64_t gdim[4] = {(cuuint64_t)d, (cuuint64_t)L, (cuuint64_t)Hkv, (cuuint64_t)B};
  cuuint64_t gstr[3] = {(cuuint64_t)k_sl * 2, (cuuint64_t)k_sh * 2, (cuuint64_t)k_sb * 2};
  if (Hkv == 1) gstr[1] = gstr[0] * (cuuint64_t)L;   // size-1 dims may carry arbitrary strides
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
  { const char* e = getenv("KVC_TC_DEBUG"); sp.dbg = e ? atoi(e) : 0; }
  sp.stats = (float2*)workspace;
  sp.out = (uint16_t*)out_scores;
  const int units = B * Hkv;
  if (dtype == KVC_BF16)
    return pl.rb == 1 ? dispatch_w<Bf16, 1>(tmap, sp, units, stream) : dispatch_w<Bf16, 2>(tmap, sp, units, stream);
  return pl.rb == 1 ? dispatch_w<Fp16, 1>(tmap, sp, units, stream) : dispatch_w<Fp16, 2>(tmap, sp, units, stream);
}

}  // namespace kvc

// Debug only: copy the clock64 trace of CTA (0,0) of the last traced launch (KVC_TC_DEBUG & 4).
extern "C" int kvc_debug_trace(long long* out_host_512) {
  return (int)cudaMemcpyFromSymbol(out_host_512, kvc::tc::g_trace, 512 * sizeof(long long));
}
