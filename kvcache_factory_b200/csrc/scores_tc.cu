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

#include "common.cuh"

namespace kvc {
namespace tc {

constexpr int kTile = 128;                    // keys per tile
constexpr int kD = 128;                       // head_dim
constexpr int kHalfBytes = kTile * 64 * 2;    // one [128 x 64] swizzled half: 16 KiB
constexpr int kTileBytes = 2 * kHalfBytes;    // 32 KiB
constexpr int kNumBuf = 4;                    // TMEM accumulator buffers (128 columns each)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (2 + kEpiWarps);

struct Params {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  int B, Hq, Hkv, L, w, groups;
  int chunk_tiles, nchunks;
  float scale;
  int reduce_mean;
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
  static constexpr int kNumBars = 2 * kStages + 2 * kNumBuf;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kCrowOff = (kTmemPtrOff + 4 + 15) & ~15;        // float[RB*128]
  static constexpr int kXstatOff = kCrowOff + RB * 128 * 4;            // float2[RB][2][128]
  static constexpr int kTotal = kXstatOff + RB * 2 * 128 * 8;
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
  const int nq = RB == 1 ? p.groups * p.w : 128;  // query rows per row block that carry data (UMMA N in pass 2)

  const uint32_t q_s = base, k_s = base + SM::kQBytes, bar_s = base + SM::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (S + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * S + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * S + kNumBuf + i); };
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::kTmemPtrOff);
  float* crow = reinterpret_cast<float*>(smem + SM::kCrowOff);
  float2* xstat = reinterpret_cast<float2*>(smem + SM::kXstatOff);

  // ---- one-time setup -------------------------------------------------------------------------
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < S; ++s) { mbar_init(kfull(s), 1); mbar_init(kempty(s), 1); }
    for (int i = 0; i < kNumBuf; ++i) { mbar_init(tfull(i), 1); mbar_init(tempty(i), 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  // Q block(s) -> shared memory in the canonical K-major SWIZZLE_128B layout:
  // byte(row, col) = half(col/64) * 16 KiB + row * 128 + ((col%64)/8 ^ row%8) * 16 + (col%8) * 2
  for (int e = tid; e < RB * 128 * 16; e += kThreads) {
    const int rg = e >> 4, c16 = e & 15;        // global stacked row, 16-byte chunk of the row
    const int rb = rg >> 7, r = rg & 127;
    const int hl = rg / W, i = rg % W;
    int4 val = make_int4(0, 0, 0, 0);
    if (hl < p.groups) {
      const uint16_t* src = p.q + b * p.q_sb + (int64_t)(kvh * p.groups + hl) * p.q_sh +
                            (int64_t)(nscore + i) * p.q_sl + c16 * 8;
      val = *reinterpret_cast<const int4*>(src);
    }
    *reinterpret_cast<int4*>(smem + rb * kTileBytes + (c16 >> 3) * kHalfBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4)) = val;
  }
  if constexpr (kPass2) {
    // fold every chunk's (m, Z) of this unit into c_row = -M*log2e - log2 Z
    for (int r = tid; r < RB * 128; r += kThreads) {
      const float2* st = p.stats + (int64_t)unit * p.nchunks * (RB * 128) + r;
      float M = -1e30f;
      for (int c = 0; c < p.nchunks; ++c) M = fmaxf(M, st[(int64_t)c * (RB * 128)].x);
      float Z = 0.f;
      for (int c = 0; c < p.nchunks; ++c) {
        const float2 s = st[(int64_t)c * (RB * 128)];
        Z += s.y * fast_exp2((s.x - M) * kLog2e);
      }
      crow[r] = (r / W < p.groups) ? (-M * kLog2e - log2f(Z)) : -INFINITY;
    }
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy Q writes -> visible to UMMA
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

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
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint32_t idesc = instr_desc(T::kDtype == KVC_BF16 ? 1 : 0, 128, kPass2 ? nq : 128);
      for (int t = 0; t < ntiles; ++t) {
        const int s = t % S;
        mbar_wait(kfull(s), (t / S) & 1);
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
    // ===== epilogue warps: two groups of four, alternating accumulator buffers =====
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may read
    const int r_loc = quad * 32 + lane;        // pass 1: query row in block; pass 2: key in tile
    float m_run[RB], z_run[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) { m_run[rb] = -1e30f; z_run[rb] = 0.f; }

    for (int j = wg; j < ntiles * RB; j += 2) {
      const int t = j / RB, rb = j % RB, buf = j % kNumBuf;
      const int key0 = (tile_begin + t) * kTile;
      mbar_wait(tfull(buf), (j / kNumBuf) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * 128;

      if constexpr (!kPass2) {
        // ---- stats: this thread owns one query row, the 128 columns are keys ----
        const int rg = rb * 128 + r_loc;
        const int hl = rg / W, wi = rg % W;
        const bool needs_mask = key0 + kTile > nscore;
        const int last_key = min(p.L - 1, hl < p.groups ? nscore + wi : p.L - 1);  // last visible key of this row
        uint32_t r2[64];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          uint32_t v[32];
          tmem_ld32(taddr + cc * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int pi = 0; pi < 16; ++pi) {
            const uint32_t p1 = T::pack(__uint_as_float(v[2 * pi]), __uint_as_float(v[2 * pi + 1]));
            uint32_t p2 = T::pack(T::lo(p1) * p.scale, T::hi(p1) * p.scale);
            if (needs_mask) {
              const int key = key0 + cc * 32 + 2 * pi;
              if (key > last_key) p2 = (p2 & 0xffff0000u) | (T::kNegInfPair & 0xffffu);
              if (key + 1 > last_key) p2 = (p2 & 0x0000ffffu) | (T::kNegInfPair & 0xffff0000u);
            }
            r2[cc * 16 + pi] = p2;
          }
        }
        // accumulator drained: hand the TMEM buffer back before the exp work
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty(buf));

        uint32_t pm = r2[0];
#pragma unroll
        for (int i = 1; i < 64; ++i) pm = T::max2(pm, r2[i]);
        const float m_new = fmaxf(m_run[rb], fmaxf(T::lo(pm), T::hi(pm)));
        float z0 = z_run[rb] * fast_exp2((m_run[rb] - m_new) * kLog2e), z1 = 0.f;
        const float ml = -m_new * kLog2e;
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          z0 += fast_exp2(fmaf(T::lo(r2[i]), kLog2e, ml));
          z1 += fast_exp2(fmaf(T::hi(r2[i]), kLog2e, ml));
        }
        m_run[rb] = m_new;
        z_run[rb] = z0 + z1;
      } else {
        // ---- scores: this thread owns one key, the columns are the stacked query rows ----
        constexpr int kHeadsPerBlk = 128 / W;
        float acc[kHeadsPerBlk];
#pragma unroll
        for (int h = 0; h < kHeadsPerBlk; ++h) acc[h] = 0.f;
        const float* cr = crow + rb * 128;
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          if (cc * 32 < nq) {
            uint32_t v[32];
            tmem_ld32(taddr + cc * 32, v);
            tmem_ld_wait();
            if ((cc + 1) * 32 >= nq) {  // last chunk read: release the accumulator buffer
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(tempty(buf));
            }
#pragma unroll
            for (int q4 = 0; q4 < 8; ++q4) {
              const float4 c4 = *reinterpret_cast<const float4*>(cr + cc * 32 + q4 * 4);
              const float cs[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
              for (int h2 = 0; h2 < 2; ++h2) {
                const int col = cc * 32 + q4 * 4 + h2 * 2;
                const uint32_t p1 = T::pack(__uint_as_float(v[q4 * 4 + h2 * 2]), __uint_as_float(v[q4 * 4 + h2 * 2 + 1]));
                const uint32_t p2 = T::pack(T::lo(p1) * p.scale, T::hi(p1) * p.scale);
                const float e0 = fast_exp2(fmaf(T::lo(p2), kLog2e, cs[h2 * 2]));
                const float e1 = fast_exp2(fmaf(T::hi(p2), kLog2e, cs[h2 * 2 + 1]));
                const uint32_t pp = T::pack(e0, e1);
                acc[col / W] += T::lo(pp) + T::hi(pp);
              }
            }
          }
        }
        const int key = key0 + r_loc;
        if (key < nscore) {
#pragma unroll
          for (int h = 0; h < kHeadsPerBlk; ++h) {
            const int head = rb * kHeadsPerBlk + h;
            if (head < p.groups) {
              float s = acc[h];
              if (p.reduce_mean) s = s / (float)p.w;
              p.out[((int64_t)b * p.Hq + kvh * p.groups + head) * nscore + key] = T::round1(s);
            }
          }
        }
      }
    }
    if constexpr (!kPass2) {
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) xstat[(rb * 2 + wg) * 128 + r_loc] = make_float2(m_run[rb], z_run[rb]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (!kPass2) {
    for (int r = tid; r < RB * 128; r += kThreads) {
      const int rb = r >> 7, rl = r & 127;
      const float2 a = xstat[(rb * 2 + 0) * 128 + rl], c = xstat[(rb * 2 + 1) * 128 + rl];
      const float mn = fmaxf(a.x, c.x);
      const float z = a.y * fast_exp2((a.x - mn) * kLog2e) + c.y * fast_exp2((c.x - mn) * kLog2e);
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
  KVC_CUDA_TRY(cudaFuncSetAttribute(k1, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  KVC_CUDA_TRY(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
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
  sp.stats = (float2*)workspace;
  sp.out = (uint16_t*)out_scores;
  const int units = B * Hkv;
  if (dtype == KVC_BF16)
    return pl.rb == 1 ? dispatch_w<Bf16, 1>(tmap, sp, units, stream) : dispatch_w<Bf16, 2>(tmap, sp, units, stream);
  return pl.rb == 1 ? dispatch_w<Fp16, 1>(tmap, sp, units, stream) : dispatch_w<Fp16, 2>(tmap, sp, units, stream);
}

}  // namespace kvc
