// Window scoring, dynamically scheduled: both softmax passes in ONE persistent kernel WITHOUT a grid barrier and
// without any wait on the critical path.
//
//   unit   = one (batch, KV head, block of 128 stacked query rows): Llama-3-8B (4 heads x w = 32) has one block per
//            KV head, Llama-3-70B (8 heads) two -- the blocks of a KV head are separate units that read the same keys
//   CTA c  serves TWO units, x = c % |X| from the first half of the units and y from the second half, in the order
//               x.pass1   y.pass1   [leftovers]   x.pass2   y.pass2
//            so x's statistics have been complete for a whole y.pass1 when x.pass2 starts, and y's for a whole x.pass2
//            when y.pass2 starts: the skew between the CTAs of a unit (measured: 5 us at cfg-2, HBM latency differs
//            per SM) is never waited for.  With a single unit the order is x.pass1, wait, x.pass2.
//   pass 1 = the unit's key tiles are cut into as many fixed chunks as CTAs serve it; chunks are CLAIMED from a
//            per-unit counter.  Every chunk yields one (m, Z) partial per query row, computed from scratch, so the
//            partials -- and with them every score bit -- do not depend on which CTA did which chunk.
//   fold   : done[unit] counts published chunks; when it is complete every CTA folds the partials in chunk order.
//   pass 2 = key tiles claimed two at a time from a second per-unit counter (descending: the most recently read tiles,
//            still in L2, first; claims are prefetched, an L2 atomic costs ~1 us); scores written.
//
// Forward progress never depends on co-residency: a CTA claims one chunk per unit in its regular order and takes
// LEFTOVER chunks (whose natural owner has not started yet: another kernel still holds its SM) only when it has
// nothing else to do; it only ever waits for chunks that were claimed by running CTAs, and a running CTA finishes
// its chunk without waiting for anybody.  No cooperative launch, any number of concurrent launches.  This replaces
// the chunked kernel + grid barrier of scores_fused.cu (round 1), whose 144 CTAs had to be co-resident.
//
// Warp roles: warp 0 = claimer + TMA producer, warp 1 = tcgen05.mma issuer, warps 2..17 = epilogue.  What a tile is
// (unit visit, pass, tile index, end of chunk) travels from the producer to the MMA issuer and the epilogue through a
// small shared-memory ring ordered by the same mbarriers that hand over the K stages and the TMEM accumulators.
// Math, rounding points and operand roles are those of scores_tc.cu / scores_fused.cu.
#include <cstdlib>
#include <mutex>

#include "tc_helpers.cuh"

namespace kvc {
namespace dyn {

using namespace tcx;

constexpr int kNumBuf = 4;
constexpr int kStages = 4;
constexpr int kEpiWarps = 16;
constexpr int kNumWG = kEpiWarps / 4;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kRing = 16;                 // > K stages + accumulator buffers: an entry is never overwritten while unread
constexpr int kC2 = 2;                    // pass-2 tiles per claim

// ring entry: code = tile | visit << 26 | pass << 27 | last-tile-of-chunk << 28, -1 = stop; aux = chunk id
constexpr int kVisitBit = 1 << 26, kPassBit = 1 << 27, kLastBit = 1 << 28, kTileMask = (1 << 26) - 1;

constexpr int kMaxUnits = 320;            // units <= 2 * #SMs
constexpr int kSlots = 16;                // concurrent launches (one control block per stream)

// Per-launch control block (library-owned, zero at module load, reset by the last CTA to exit).
struct Ctl {
  unsigned p1_next[kMaxUnits];   // next pass-1 chunk of the unit
  unsigned p2_next[kMaxUnits];   // next pass-2 claim (kC2 tiles each, counted from the end)
  unsigned done[kMaxUnits];      // pass-1 chunks whose statistics are published
  unsigned exited;
};
__device__ Ctl g_ctl[kSlots];

struct Params {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  int B, Hq, Hkv, L, w, groups;
  int nblk;             // 128-row blocks per KV head
  int units, nx, ny;    // units = B * Hkv * nblk; |X| = nx, |Y| = ny
  int max_chunks;       // partial slots per unit in `stats`
  float scale;
  int reduce_mean;
  int stagger_ns;       // see the epilogue: nanoseconds between the warpgroups' restarts after a CTA-wide sync
  float2* stats;        // [units][max_chunks][128]
  const float* lse;     // optional [B,Hq,w]: log-sum-exp of every window query row over all keys (natural log, scaled
                        // logits), e.g. the softmax_lse the prefill's flash attention returns: pass 1 is skipped
  uint16_t* out;        // [B,Hq,L-w]
  unsigned long long* tl;
};

struct Smem {
  static constexpr int kQBytes = 2 * kTileBytes;                       // the Q blocks of the two units
  static constexpr int kKBytes = kStages * kTileBytes;
  static constexpr int kBarOff = kQBytes + kKBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kNumBuf + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kRingOff = (kTmemPtrOff + 8 + 15) & ~15;        // int2[kRing]
  static constexpr int kCrowOff = kRingOff + kRing * 8;                // float[2][128]
  static constexpr int kXstatOff = kCrowOff + 2 * 128 * 4;             // float2[kNumWG][128]
  static constexpr int kTotal = kXstatOff + kNumWG * 128 * 8;
  static constexpr int kAlloc = kTotal + 1024;
};

// Developer timeline (`make TRACE=1` builds libkvc_b200_trace.so): 8 stamps per CTA, read back by
// kvc_debug_dyn_trace().  Compiled out of the shipped library.
#ifdef KVC_TRACE
__device__ long long g_dtrace[160 * 8];
__device__ __forceinline__ void dstamp(int slot, long long v = -1) {
  long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  g_dtrace[blockIdx.x * 8 + slot] = v >= 0 ? v : t;
}
#define DYN_STAMP(cond, ...) do { if (cond) dstamp(__VA_ARGS__); } while (0)
// CTA 0, the four epilogue warps of SM sub-partition 2 (warps 2, 6, 10, 14: one per warpgroup), first 64 tiles:
// SM-clock stamps {accumulator seen, columns loaded, math done, ring code}
__device__ unsigned g_wtrace[4][64][4];
__device__ __forceinline__ void wstamp(int wg, unsigned ts, int slot, unsigned v = 0xffffffffu) {
  if (blockIdx.x == 0 && ts < 64) g_wtrace[wg][ts][slot] = v != 0xffffffffu ? v : (unsigned)clock64();
}
#define DYN_WSTAMP(...) do { if ((threadIdx.x & 127) == 64 && (threadIdx.x >> 5) >= 2) wstamp(__VA_ARGS__); } while (0)
#else
#define DYN_STAMP(cond, ...) do { } while (0)
#define DYN_WSTAMP(...) do { } while (0)
#endif

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

// CTAs of a grid of G that serve unit u (u < nx: c % nx == u; else c % ny == u - nx), and the chunking of its tiles.
__device__ __host__ __forceinline__ int serving_ctas(int u, int nx, int ny, int G) {
  const int m = u < nx ? nx : ny, r = u < nx ? u : u - nx;
  return r < G ? (G - 1 - r) / m + 1 : 0;
}

template <typename T, int W, int POLY1, int POLY2>
__global__ void __launch_bounds__(kThreads, 1)
window_score_dyn_kernel(const __grid_constant__ CUtensorMap tmap_k, const Params p, Ctl* __restrict__ ctl) {
  using SM = Smem;
  constexpr int S = kStages;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int G = gridDim.x;
  const int nvis = p.ny > 0 ? 2 : 1;
  // the two units of this CTA: unit -> (batch, KV head, row block)
  // (two scalars + select instead of arrays indexed by the visit: dynamic indexing would put them in local memory)
  const int unit_0 = blockIdx.x % p.nx;
  const int unit_1 = p.ny > 0 ? p.nx + blockIdx.x % p.ny : unit_0;
  const int nscore = p.L - p.w;
  const int T1 = (p.L + kTile - 1) / kTile;        // pass-1 tiles of a unit
  const int T2 = (nscore + kTile - 1) / kTile;     // pass-2 tiles (tiles that hold only window keys are skipped)
  const int ublk_0 = unit_0 % p.nblk, ublk_1 = unit_1 % p.nblk;
  const int ukvh_0 = (unit_0 / p.nblk) % p.Hkv, ukvh_1 = (unit_1 / p.nblk) % p.Hkv;
  const int ub_0 = unit_0 / (p.nblk * p.Hkv), ub_1 = unit_1 / (p.nblk * p.Hkv);
  const int n1c_0 = min(serving_ctas(unit_0, p.nx, p.ny, G), T1);   // pass-1 chunks of the unit
  const int n1c_1 = min(serving_ctas(unit_1, p.nx, p.ny, G), T1);
#define VSEL(name, v) ((v) ? name##_1 : name##_0)

  const uint32_t q_s = base, k_s = base + SM::kQBytes, bar_s = base + SM::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (S + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * S + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * S + kNumBuf + i); };
  const uint32_t qbar = bar_s + 8u * (2 * S + 2 * kNumBuf);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::kTmemPtrOff);
  volatile int2* ring = reinterpret_cast<volatile int2*>(smem + SM::kRingOff);
  float* crow = reinterpret_cast<float*>(smem + SM::kCrowOff);
  float2* xstat = reinterpret_cast<float2*>(smem + SM::kXstatOff);

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
  // Programmatic dependent launch: the prologue above may overlap the tail of the previous kernel in the stream;
  // nothing below may start before that kernel has completed (inputs, workspace, the control block).
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

  if (warp == 0) {
    // ===== claimer + TMA producer =====
    // The whole warp runs the loops (uniform control flow keeps addresses in uniform registers); lane 0 claims, one
    // elected lane issues.
    int s = 0, ph = 1, ts = 0;
    auto emit = [&](int v, int pass, int tile, bool last, int chunk) {
      mbar_wait(kempty(s), ph);
      if (elect_one()) {
        ring[ts % kRing].y = chunk;
        ring[ts % kRing].x = tile | (v ? kVisitBit : 0) | (pass ? kPassBit : 0) | (last ? kLastBit : 0);
        mbar_expect_tx(kfull(s), kTileBytes);
        tma_load_4d(k_s + s * kTileBytes, &tmap_k, kfull(s), 0, tile * kTile, VSEL(ukvh, v), VSEL(ub, v));
        tma_load_4d(k_s + s * kTileBytes + kHalfBytes, &tmap_k, kfull(s), 64, tile * kTile, VSEL(ukvh, v), VSEL(ub, v));
      }
      __syncwarp();
      ++ts;
      if (++s == S) { s = 0; ph ^= 1; }
    };
    auto emit_chunk = [&](int v, int j) {
      const int t0 = (int)(((long long)j * T1) / VSEL(n1c, v)), t1 = (int)(((long long)(j + 1) * T1) / VSEL(n1c, v));
      for (int t = t0; t < t1; ++t) emit(v, 0, t, t == t1 - 1, j);
    };
    // regular pass-1 work: one chunk of x, one of y (both claims in flight together)
    unsigned c0 = 0xffffffffu, c1 = 0xffffffffu;
    if (lane == 0 && p.lse == nullptr) {
      c0 = atomicAdd(&ctl->p1_next[unit_0], 1u);
      if (nvis == 2) c1 = atomicAdd(&ctl->p1_next[unit_1], 1u);
    }
    c0 = __shfl_sync(0xffffffffu, c0, 0);
    c1 = __shfl_sync(0xffffffffu, c1, 0);
    if (c0 < (unsigned)n1c_0) emit_chunk(0, (int)c0);
    if (nvis == 2 && c1 < (unsigned)n1c_1) emit_chunk(1, (int)c1);
    // leftovers (chunks whose natural owner is not running yet) + the first pass-2 claim, all in flight together
    unsigned p2c = 0;
    for (; p.lse == nullptr;) {
      unsigned l0 = 0, l1 = 0;
      if (lane == 0) {
        l0 = atomicAdd(&ctl->p1_next[unit_0], 1u);
        if (nvis == 2) l1 = atomicAdd(&ctl->p1_next[unit_1], 1u);
      }
      l0 = __shfl_sync(0xffffffffu, l0, 0);
      l1 = __shfl_sync(0xffffffffu, l1, 0);
      const bool got0 = (int)l0 < n1c_0, got1 = nvis == 2 && (int)l1 < n1c_1;
      if (got0) emit_chunk(0, (int)l0);
      if (got1) emit_chunk(1, (int)l1);
      if (!got0 && !got1) break;
    }
    // pass 2 of x, then of y: kC2 tiles per claim, the next claim always in flight
    for (int v = 0; v < nvis; ++v) {
      unsigned* ctr = &ctl->p2_next[VSEL(unit, v)];
      if (lane == 0) p2c = atomicAdd(ctr, 1u);
      for (;;) {
        const unsigned k = __shfl_sync(0xffffffffu, p2c, 0);
        const int first = (int)k * kC2;
        if (first >= T2) break;
        if (lane == 0) p2c = atomicAdd(ctr, 1u);
#pragma unroll
        for (int i = 0; i < kC2; ++i)
          if (first + i < T2) emit(v, 1, T2 - 1 - (first + i), false, 0);
      }
    }
    mbar_wait(kempty(s), ph);
    if (elect_one()) { ring[ts % kRing].x = -1; mbar_arrive(kfull(s)); }
  } else if (warp == 1) {
    // ===== MMA issuer: pass 1 D = Q.K^T (lanes = query rows), pass 2 D = K.Q^T (lanes = keys) =====
    const int fmt = T::kDtype == KVC_BF16 ? 1 : 0;
    const uint32_t idesc2_0 = instr_desc(fmt, 128, min(128, p.groups * p.w - ublk_0 * 128));
    const uint32_t idesc2_1 = instr_desc(fmt, 128, min(128, p.groups * p.w - ublk_1 * 128));
    const uint32_t idesc1 = instr_desc(fmt, 128, 128);
    mbar_wait(qbar, 0);
    tc_fence_after();
    int s = 0, ph = 0, buf = 0, bph = 1;
    for (int ts = 0;; ++ts) {
      mbar_wait(kfull(s), ph);
      const int code = ring[ts % kRing].x;
      mbar_wait(tempty(buf), bph);
      if (code < 0) {   // stop: hand the marker on to the epilogue through the next accumulator barrier
        if (elect_one()) mbar_arrive(tfull(buf));
        break;
      }
      tc_fence_after();
      const bool second = (code & kPassBit) != 0;
      const int v = (code & kVisitBit) ? 1 : 0;
      const uint32_t idesc = second ? VSEL(idesc2, v) : idesc1;
      const uint64_t kd0 = smem_desc(k_s + s * kTileBytes), qd0 = smem_desc(q_s + v * kTileBytes);
      const uint64_t ad0 = second ? kd0 : qd0, bd0 = second ? qd0 : kd0;
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t off = (uint64_t)(((k >> 2) * kHalfBytes + (k & 3) * 32) >> 4);  // start-address field is in 16-byte units
          tc_mma(tmem_base + buf * 128, ad0 + off, bd0 + off, idesc, k > 0);
        }
        tc_commit(tfull(buf));
        tc_commit(kempty(s));
      }
      __syncwarp();
      if (++buf == kNumBuf) { buf = 0; bph ^= 1; }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    // ===== epilogue warps: warpgroup g owns columns [32g, 32g+32) of every accumulator tile =====
    const int et = tid - 64;
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r_loc = quad * 32 + lane;
    const float scale = p.scale;
    // stage the Q blocks of both units: canonical K-major SWIZZLE_128B layout; rows beyond the last head are zero
#pragma unroll
    for (int it = 0; it < (2 * 128 * 16) / kEpiThreads; ++it) {
      const int e = et + it * kEpiThreads;
      const int v = e >> 11, r = (e >> 4) & 127, c16 = e & 15;
      const int rg = VSEL(ublk, v) * 128 + r;
      const int hl = rg / W, i = rg % W;
      const bool ok = hl < p.groups && v < nvis;
      const uint16_t* src = p.q + VSEL(ub, v) * p.q_sb + (int64_t)(VSEL(ukvh, v) * p.groups + (ok ? hl : 0)) * p.q_sh +
                            (int64_t)(nscore + i) * p.q_sl + c16 * 8;
      const uint32_t dst = q_s + v * kTileBytes + (c16 >> 3) * kHalfBytes + r * 128 + (((c16 & 7) ^ (r & 7)) << 4);
      const int sz = ok ? 16 : 0;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_all;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_arrive(qbar);
    DYN_STAMP(et == 0, 0);

    // running statistics of the pass-1 chunk in flight (one chunk at a time)
    float m_run = -1e30f;
    float2 zacc[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
    const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + wg * 32;
    // rows beyond the unit's last head are zero padding: warps that own only padding rows skip the load and the exps
    const bool dead_rows_0 = __all_sync(0xffffffffu, (ublk_0 * 128 + r_loc) / W >= p.groups);
    const bool dead_rows_1 = __all_sync(0xffffffffu, (ublk_1 * 128 + r_loc) / W >= p.groups);
    constexpr int kHeadsPerChunk = 32 / W;
    bool folded_0 = false, folded_1 = false;
    int n1_tiles = 0, n2_tiles = 0, npub = 0;
    bool restagger = true;

    for (unsigned ts = 0;; ++ts) {
      const int buf = ts % kNumBuf;
      mbar_wait(tfull(buf), (ts / kNumBuf) & 1);
      tc_fence_after();
      const int code = ring[ts % kRing].x;
      DYN_STAMP(et == 0 && ts == 0, 1);
      DYN_WSTAMP(wg, ts, 0);
      DYN_WSTAMP(wg, ts, 3, (unsigned)code);
      if (restagger) {
        restagger = false;
        if (wg > 0 && p.stagger_ns > 0) __nanosleep(wg * p.stagger_ns);
      }
      if (code < 0) break;
      const int tile = code & kTileMask;
      const int v = (code & kVisitBit) ? 1 : 0;
      if (!(code & kPassBit)) {
        // ---------------- pass 1: thread <-> query row, columns are keys ----------------
        const int key0 = tile * kTile + wg * 32;
        if (VSEL(dead_rows, v)) {
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
        } else {
          uint32_t vals[32];
          tmem_ld32(tcol + buf * 128, vals);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
          DYN_WSTAMP(wg, ts, 1);
          if (key0 + 32 > nscore) {  // chunk touches the w x w causal block or the ragged end
            const int rg = VSEL(ublk, v) * 128 + r_loc;
            const int last_key = min(p.L - 1, rg / W < p.groups ? nscore + rg % W : p.L - 1);
            p1_chunk<T, true, POLY1>(vals, scale, m_run, zacc, key0, last_key);
          } else {
            p1_chunk<T, false, POLY1>(vals, scale, m_run, zacc, 0, 0);
          }
        }
        ++n1_tiles;
        if (zacc[0].x >= 0.f) DYN_WSTAMP(wg, ts, 2);
        if (code & kLastBit) {
          // ---- end of a chunk: combine the four column slices (fixed order), publish the partial ----
          const int chunk = ring[ts % kRing].y;
          const float2 z2 = __fadd2_rn(__fadd2_rn(zacc[0], zacc[1]), __fadd2_rn(zacc[2], zacc[3]));
          xstat[wg * 128 + r_loc] = make_float2(m_run, z2.x + z2.y);
          epi_sync();
          if (et < 128) {
            float mn = -1e30f, z = 0.f;
#pragma unroll
            for (int g = 0; g < kNumWG; ++g) mn = fmaxf(mn, xstat[g * 128 + et].x);
#pragma unroll
            for (int g = 0; g < kNumWG; ++g) {
              const float2 a = xstat[g * 128 + et];
              z += a.y * fast_exp2((a.x - mn) * kLog2e);
            }
            p.stats[((int64_t)VSEL(unit, v) * p.max_chunks + chunk) * 128 + et] = make_float2(mn, z);
            __threadfence();
          }
          epi_sync();
          if (et == 0) atomicAdd(&ctl->done[VSEL(unit, v)], 1u);
          m_run = -1e30f;
#pragma unroll
          for (int a = 0; a < 4; ++a) zacc[a] = make_float2(0.f, 0.f);
          DYN_STAMP(et == 0 && npub < 2, 2 + npub);
          ++npub;
          restagger = true;
        }
      } else {
        // ---------------- pass 2: thread <-> key, columns are the stacked query rows ----------------
        if (!VSEL(folded, v)) {
          // Fold the unit's partials (chunk order) into c_row = -M log2e - log2 Z.  Every chunk of the unit was
          // claimed by a RUNNING CTA before any pass-2 tile of it could be claimed, and those CTAs publish without
          // waiting for anybody: this wait always ends (normally it does not wait at all, see the header).
          if (v) folded_1 = true; else folded_0 = true;
          if (p.lse != nullptr) {
            // statistics supplied by the caller: c_row = -lse * log2(e); no pass 1, nothing to wait for
            if (et < 128) {
              const int rg = VSEL(ublk, v) * 128 + et, hl = rg / W, i = rg % W;
              crow[v * 128 + et] = hl < p.groups
                  ? -p.lse[((int64_t)VSEL(ub, v) * p.Hq + VSEL(ukvh, v) * p.groups + hl) * p.w + i] * kLog2e : -INFINITY;
            }
            epi_sync();
          } else {
          if (et == 0) {
            for (unsigned spin = 0; ld_acquire_u32(&ctl->done[VSEL(unit, v)]) < (unsigned)VSEL(n1c, v); ++spin) {
              __nanosleep(40);
              if (spin > (1u << 24)) {
                printf("kvc scores_dyn: statistics of unit %d never completed (block %d)\n", VSEL(unit, v), blockIdx.x);
                __trap();
              }
            }
          }
          epi_sync();
          const int nparts = VSEL(n1c, v);
          constexpr int kParts = kEpiThreads / 128;
          const int r = et % 128, q = et / 128;
          const float2* st = p.stats + (int64_t)VSEL(unit, v) * p.max_chunks * 128 + r;
          float M = -1e30f, Z = 0.f;
          for (int c0 = q; c0 < nparts; c0 += 8 * kParts) {
            float2 sv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              sv[u] = make_float2(-1e30f, 0.f);
              if (c0 + u * kParts < nparts) {
                const float2* a = st + (int64_t)(c0 + u * kParts) * 128;
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
          xstat[q * 128 + r] = make_float2(M, Z);
          epi_sync();
          if (et < 128) {
            float mx = -1e30f, z = 0.f;
#pragma unroll
            for (int g = 0; g < kParts; ++g) mx = fmaxf(mx, xstat[g * 128 + et].x);
#pragma unroll
            for (int g = 0; g < kParts; ++g) {
              const float2 a = xstat[g * 128 + et];
              z += a.y * fast_exp2((a.x - mx) * kLog2e);
            }
            crow[v * 128 + et] = ((VSEL(ublk, v) * 128 + et) / W < p.groups) ? (-mx * kLog2e - log2f(z)) : -INFINITY;
          }
          epi_sync();
          }
          DYN_STAMP(et == 0, 4 + v);
          if (wg > 0 && p.stagger_ns > 0) __nanosleep(wg * p.stagger_ns);
        }
        const int key = tile * kTile + r_loc;
        const int head0 = (VSEL(ublk, v) * 128 + wg * 32) / W;
        if (head0 >= p.groups) {  // this warpgroup's 32 columns are beyond the last head
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
        } else {
          uint32_t vals[32];
          tmem_ld32(tcol + buf * 128, vals);
          tmem_ld_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(tempty(buf));
          DYN_WSTAMP(wg, ts, 1);
          float2 acc[kHeadsPerChunk];
#pragma unroll
          for (int h = 0; h < kHeadsPerChunk; ++h) acc[h] = make_float2(0.f, 0.f);
          p2_chunk<T, W, POLY2>(vals, scale, crow + v * 128 + wg * 32, acc);
          if (key < nscore) {
#pragma unroll
            for (int h = 0; h < kHeadsPerChunk; ++h) {
              const int head = head0 + h;
              if (head < p.groups) {
                float sum = acc[h].x + acc[h].y;
                if (p.reduce_mean) sum = div_exact(sum, (float)p.w, 1.0f / (float)p.w);
                p.out[((int64_t)VSEL(ub, v) * p.Hq + VSEL(ukvh, v) * p.groups + head) * nscore + key] = T::round1(sum);
              }
            }
          }
        }
        ++n2_tiles;
        DYN_WSTAMP(wg, ts, 2);
      }
    }
    DYN_STAMP(et == 0, 6);
    DYN_STAMP(et == 0, 7, (long long)(n1_tiles * 1000 + n2_tiles));
  }

  tc_fence_before();
  __syncthreads();
  tl_end(p.tl);
  if (tid == 0) {
    // the last CTA to leave resets the control block for the next launch on this stream
    __threadfence();
    const unsigned old = atomicAdd(&ctl->exited, 1u);
    if (old == gridDim.x - 1) {
      for (int u = 0; u < p.units; ++u) { ctl->p1_next[u] = 0u; ctl->p2_next[u] = 0u; ctl->done[u] = 0u; }
      ctl->exited = 0u;
      __threadfence();
    }
  }
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// Control block of a stream: launches on one stream never overlap (the kernel starts with griddepcontrol.wait),
// launches on different streams must not share counters.  More than kSlots distinct streams per device are refused.
static int ctl_slot(cudaStream_t stream) {
  static std::mutex mu;
  static cudaStream_t seen[64][kSlots];
  static int n[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return -1;
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n[dev]; ++i)
    if (seen[dev][i] == stream) return i;
  if (n[dev] == kSlots) return -1;
  seen[dev][n[dev]] = stream;
  return n[dev]++;
}

struct Plan {
  int nblk, units, nx, ny, grid, max_chunks;
  size_t stats_bytes;
};

static Plan make_plan(int B, int Hq, int Hkv, int L, int w) {
  Plan pl{};
  const int g = Hq / Hkv;
  pl.nblk = (g * w + 127) / 128;
  pl.units = B * Hkv * pl.nblk;
  pl.nx = (pl.units + 1) / 2;
  pl.ny = pl.units - pl.nx;
  const int tiles = (L + kTile - 1) / kTile;
  const int sms = sm_count();
  const long long want = (long long)pl.units * tiles;      // never more CTAs than pass-1 tiles
  pl.grid = (int)(want < sms ? want : sms);
  if (pl.grid < pl.nx) pl.grid = pl.nx;                     // every unit is served by at least one CTA
  pl.max_chunks = 0;
  for (int u = 0; u < pl.units; ++u) {
    const int n = serving_ctas(u, pl.nx, pl.ny, pl.grid);
    if (n > pl.max_chunks) pl.max_chunks = n;
  }
  pl.stats_bytes = ((size_t)pl.units * pl.max_chunks * 128 * sizeof(float2) + 255) & ~(size_t)255;
  return pl;
}

template <typename T, int W>
static int launch(const CUtensorMap& tmap, const Params& sp, const Plan& pl, cudaStream_t stream) {
  // Exponentials on the FMA pipe: 4 of 16 pairs in each pass (scripts/ubench/epi_mix.cu: 1086 -> 881 and
  // 1088 -> 939 cycles per tile of pure epilogue math); KVC_DYN_POLY=0 keeps every exponential on MUFU.
  static const bool poly = [] { const char* e = getenv("KVC_DYN_POLY"); return !(e && e[0] == '0'); }();
  auto kern = poly ? window_score_dyn_kernel<T, W, 4, 4> : window_score_dyn_kernel<T, W, 0, 0>;
  const int smem = Smem::kAlloc;
  KVC_SET_MAX_SMEM(kern, smem);
  const int slot = ctl_slot(stream);
  KVC_REQUIRE(slot >= 0, KVC_ERR_UNSUPPORTED, "dynamic tcgen05 scoring kernel: more than %d distinct streams on this device", kSlots);
  Ctl* ctl = nullptr;
  KVC_CUDA_TRY(cudaGetSymbolAddress((void**)&ctl, g_ctl));
  ctl += slot;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(pl.grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  static const bool pdl = [] { const char* e = getenv("KVC_PDL"); return !(e && e[0] == '0'); }();
  cfg.numAttrs = pdl ? 1 : 0;
  KVC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, tmap, sp, ctl));
  return launch_status("window_score_dyn_kernel");
}

template <typename T>
static int dispatch_w(const CUtensorMap& tmap, const Params& sp, const Plan& pl, cudaStream_t stream) {
  switch (sp.w) {
    case 8: return launch<T, 8>(tmap, sp, pl, stream);
    case 16: return launch<T, 16>(tmap, sp, pl, stream);
    default: return launch<T, 32>(tmap, sp, pl, stream);
  }
}

}  // namespace dyn

#ifdef KVC_TRACE
extern "C" int kvc_debug_dyn_trace(long long* out, int n) {
  return (int)cudaMemcpyFromSymbol(out, dyn::g_dtrace, sizeof(long long) * (size_t)n);
}
extern "C" int kvc_debug_dyn_warp_trace(unsigned* out) {
  return (int)cudaMemcpyFromSymbol(out, dyn::g_wtrace, sizeof(unsigned) * 4 * 64 * 4);
}
#endif

bool window_scores_tc_supported(int B, int Hq, int Hkv, int L, int d, int w);

// Every unit (batch, KV head, 128-row block) needs at least one CTA in each half: units <= 2 * #SMs.
bool window_scores_dyn_supported(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_tc_supported(B, Hq, Hkv, L, d, w)) return false;
  if (w > 32) return false;  // a 64-row head spans two warpgroups' columns: the two-launch kernel covers it
  const int g = Hq / Hkv;
  if (g * w > 128 && (g * w) % 16) return false;            // pass-2 N of a partial row block must be a multiple of 16
  const long long units = (long long)B * Hkv * ((g * w + 127) / 128);
  return (units + 1) / 2 <= sm_count() && units <= dyn::kMaxUnits;
}

size_t window_scores_dyn_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_dyn_supported(B, Hq, Hkv, L, d, w)) return 0;
  return dyn::make_plan(B, Hq, Hkv, L, w).stats_bytes + 256;
}

int window_scores_dyn(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                      int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce,
                      void* out_scores, void* workspace, size_t workspace_bytes, const float* lse, cudaStream_t stream) {
  using namespace dyn;
  KVC_REQUIRE(window_scores_dyn_supported(B, Hq, Hkv, L, d, w), KVC_ERR_UNSUPPORTED,
              "dynamic tcgen05 scoring kernel does not cover this shape (Hq/Hkv=%d, w=%d, d=%d, B*Hkv=%d)", Hq / Hkv, w, d,
              B * Hkv);
  KVC_REQUIRE(((q_sb | q_sh | q_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel: query rows must be 16-byte aligned");
  KVC_REQUIRE(((k_sb | k_sh | k_sl) & 7) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0, KVC_ERR_UNSUPPORTED,
              "tcgen05 scoring kernel: key rows must be 16-byte aligned (TMA)");
  const Plan pl = make_plan(B, Hq, Hkv, L, w);
  KVC_REQUIRE(workspace_bytes >= pl.stats_bytes + 256, KVC_ERR_WORKSPACE, "window_scores workspace: have %zu need %zu",
              workspace_bytes, pl.stats_bytes + 256);
  CUtensorMap tmap;
  if (int rc = make_k_tensor_map(&tmap, k, k_sb, k_sh, k_sl, B, Hkv, L, d, dtype)) return rc;
  Params sp{};
  sp.q = (const uint16_t*)q; sp.q_sb = q_sb; sp.q_sh = q_sh; sp.q_sl = q_sl;
  sp.B = B; sp.Hq = Hq; sp.Hkv = Hkv; sp.L = L; sp.w = w; sp.groups = Hq / Hkv;
  sp.nblk = pl.nblk; sp.units = pl.units; sp.nx = pl.nx; sp.ny = pl.ny; sp.max_chunks = pl.max_chunks;
  sp.scale = 1.0f / sqrtf((float)d);
  sp.reduce_mean = reduce == KVC_REDUCE_MEAN;
  static const int stagger = [] { const char* e = getenv("KVC_DYN_STAGGER_NS"); return e ? atoi(e) : 0; }();
  sp.stagger_ns = stagger;
  sp.stats = (float2*)workspace;
  sp.lse = lse;
  sp.out = (uint16_t*)out_scores;
  sp.tl = KVC_TL_SLOT(0);
  return dtype == KVC_BF16 ? dispatch_w<Bf16>(tmap, sp, pl, stream) : dispatch_w<Fp16>(tmap, sp, pl, stream);
}

}  // namespace kvc
