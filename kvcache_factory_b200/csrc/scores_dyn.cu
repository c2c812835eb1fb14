// Window scoring, dynamically scheduled: both softmax passes in ONE persistent kernel WITHOUT a grid barrier.
//
//   unit   = one (batch, KV head) with its g query heads stacked into RB blocks of 128 query rows
//   CTA c  serves unit c % units for its whole life (the unit's Q block is staged once)
//   pass 1 = the unit's key tiles are CLAIMED one by one from a per-unit counter (ascending);
//            every CTA keeps running (m, Z) per query row over the tiles it got
//   publish: (m, Z) partials -> global; done[unit] += tiles processed
//   wait   : until done[unit] == all pass-1 tiles of the unit (a count of FINISHED TILES, not of CTAs)
//   fold   : every CTA folds the unit's partials into c_row = -M log2e - log2 Z
//   pass 2 = key tiles claimed from a second per-unit counter (descending: the most recently read tiles, still
//            in L2, first); scores written
//
// Why this replaces the chunked kernel + grid barrier of scores_fused.cu (round 1):
//   * forward progress never depends on co-residency.  A CTA only ever waits for tiles that were claimed by CTAs
//     that are running, and a running CTA never waits before it has finished its own pass-1 tiles.  A CTA that starts
//     late (other kernels still hold its SM, a second launch on another stream) simply finds fewer tiles left.  No
//     cooperative launch, any number of concurrent launches.
//   * tile-granular balance over ALL SMs (148 CTAs; the chunked kernel ran 144 CTAs with 15/15/.../1 tiles), and the
//     wait is per unit: only the CTAs that share a softmax row meet.
//   * the TMA producer and the MMA issuer never wait for the statistics: when the epilogue reaches pass 2 the K ring
//     and the TMEM accumulators already hold pass-2 tiles.
//
// Warp roles: warp 0 = tile claimer + TMA producer, warp 1 = tcgen05.mma issuer, warps 2.. = epilogue.  The tile
// index travels from the producer to the MMA issuer and the epilogue through a small shared-memory ring
// (entry = tile | pass flag, -1 = stop) ordered by the same mbarriers that hand over the K stages and accumulators.
// Math, rounding points and operand roles are those of scores_tc.cu / scores_fused.cu.
#include <cstdlib>
#include <mutex>

#include "tc_helpers.cuh"

namespace kvc {
namespace dyn {

using namespace tcx;

constexpr int kNumBuf = 4;
constexpr int kEpiWarps = 16;
constexpr int kNumWG = kEpiWarps / 4;
constexpr int kEpiThreads = kEpiWarps * 32;
constexpr int kThreads = 32 * (2 + kEpiWarps);
constexpr int kRing = 16;                 // > K stages + accumulator buffers: an entry is never overwritten while unread
constexpr int kP2Flag = 1 << 30;

constexpr int kMaxUnits = 160;            // units <= #SMs (148)
constexpr int kSlots = 16;                // concurrent launches (one control block per stream)

// Per-launch control block (library-owned, zero at module load, reset by the last CTA to exit).
struct Ctl {
  unsigned p1_next[kMaxUnits];   // next pass-1 tile of the unit
  unsigned p2_next[kMaxUnits];   // next pass-2 tile (counted from the end)
  unsigned done[kMaxUnits];      // pass-1 tiles whose statistics are published
  unsigned nparts[kMaxUnits];    // partial slots handed out
  unsigned exited;
};
__device__ Ctl g_ctl[kSlots];

struct Params {
  const uint16_t* q; int64_t q_sb, q_sh, q_sl;
  int B, Hq, Hkv, L, w, groups;
  int units, slots_per_unit;
  float scale;
  int reduce_mean;
  float2* stats;        // [units][slots_per_unit][RB*128]
  uint16_t* out;        // [B,Hq,L-w]
  unsigned long long* tl;
};

template <int RB>
struct Smem {
  static constexpr int kStages = RB == 1 ? 5 : 4;
  static constexpr int kQBytes = RB * kTileBytes;
  static constexpr int kKBytes = kStages * kTileBytes;
  static constexpr int kBarOff = kQBytes + kKBytes;
  static constexpr int kNumBars = 2 * kStages + 2 * kNumBuf + 1;
  static constexpr int kTmemPtrOff = kBarOff + kNumBars * 8;
  static constexpr int kRingOff = (kTmemPtrOff + 8 + 15) & ~15;        // int[kRing]
  static constexpr int kMiscOff = kRingOff + kRing * 4;                // int[4]: partial slot broadcast
  static constexpr int kCrowOff = kMiscOff + 16;                       // float[RB*128]
  static constexpr int kXstatOff = kCrowOff + RB * 128 * 4;            // float2[RB][kNumWG][128]
  static constexpr int kTotal = kXstatOff + RB * kNumWG * 128 * 8;
  static constexpr int kAlloc = kTotal + 1024;
};

__device__ __forceinline__ void epi_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }

template <typename T, int RB, int W, int POLY1, int POLY2>
__global__ void __launch_bounds__(kThreads, 1)
window_score_dyn_kernel(const __grid_constant__ CUtensorMap tmap_k, const Params p, Ctl* __restrict__ ctl) {
  using SM = Smem<RB>;
  constexpr int S = SM::kStages;
  constexpr int kRows = RB * 128;
  extern __shared__ unsigned char smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  const uint32_t base = (raw_addr + 1023u) & ~1023u;
  unsigned char* smem = smem_raw + (base - raw_addr);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int unit = blockIdx.x % p.units;
  const int b = unit / p.Hkv, kvh = unit % p.Hkv;
  const int nscore = p.L - p.w;
  const int T1 = (p.L + kTile - 1) / kTile;        // pass-1 tiles of a unit
  const int T2 = (nscore + kTile - 1) / kTile;     // pass-2 tiles (tiles that hold only window keys are skipped)
  const int nq = RB == 1 ? p.groups * p.w : 128;

  const uint32_t q_s = base, k_s = base + SM::kQBytes, bar_s = base + SM::kBarOff;
  auto kfull = [&](int s) { return bar_s + 8u * s; };
  auto kempty = [&](int s) { return bar_s + 8u * (S + s); };
  auto tfull = [&](int i) { return bar_s + 8u * (2 * S + i); };
  auto tempty = [&](int i) { return bar_s + 8u * (2 * S + kNumBuf + i); };
  const uint32_t qbar = bar_s + 8u * (2 * S + 2 * kNumBuf);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + SM::kTmemPtrOff);
  volatile int* ring = reinterpret_cast<volatile int*>(smem + SM::kRingOff);
  volatile int* misc = reinterpret_cast<volatile int*>(smem + SM::kMiscOff);
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
    // ===== tile claimer + TMA producer =====
    // The whole warp runs the loop (uniform control flow keeps addresses in uniform registers); lane 0 claims, one
    // elected lane issues.  The claim for the NEXT tile is issued before this tile's stage wait, so the atomic's
    // round trip to L2 is off the critical path.
    int s = 0, ph = 1, ts = 0, pass = 0;
    unsigned raw = 0;
    if (lane == 0) raw = atomicAdd(&ctl->p1_next[unit], 1u);
    for (;;) {
      unsigned t = __shfl_sync(0xffffffffu, raw, 0);
      if (pass == 0 && (int)t >= T1) {   // pass-1 tiles of the unit are all claimed: go on with pass 2 right away
        pass = 1;
        if (lane == 0) raw = atomicAdd(&ctl->p2_next[unit], 1u);
        t = __shfl_sync(0xffffffffu, raw, 0);
      }
      const bool stop = pass == 1 && (int)t >= T2;
      if (!stop && lane == 0) raw = atomicAdd(pass == 0 ? &ctl->p1_next[unit] : &ctl->p2_next[unit], 1u);
      mbar_wait(kempty(s), ph);
      if (stop) {
        if (elect_one()) { ring[ts % kRing] = -1; mbar_arrive(kfull(s)); }
        break;
      }
      const int tile = pass == 0 ? (int)t : T2 - 1 - (int)t;
      if (elect_one()) {
        ring[ts % kRing] = tile | (pass ? kP2Flag : 0);
        mbar_expect_tx(kfull(s), kTileBytes);
        tma_load_4d(k_s + s * kTileBytes, &tmap_k, kfull(s), 0, tile * kTile, kvh, b);
        tma_load_4d(k_s + s * kTileBytes + kHalfBytes, &tmap_k, kfull(s), 64, tile * kTile, kvh, b);
      }
      __syncwarp();
      ++ts;
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: pass 1 D = Q.K^T (lanes = query rows), pass 2 D = K.Q^T (lanes = keys) =====
    const int fmt = T::kDtype == KVC_BF16 ? 1 : 0;
    const uint32_t idesc1 = instr_desc(fmt, 128, 128), idesc2 = instr_desc(fmt, 128, nq);
    mbar_wait(qbar, 0);
    tc_fence_after();
    int s = 0, ph = 0, buf = 0, bph = 1;
    for (int ts = 0;; ++ts) {
      mbar_wait(kfull(s), ph);
      const int entry = ring[ts % kRing];
      if (entry < 0) {   // stop: hand the marker on to the epilogue through the next accumulator barrier
        mbar_wait(tempty(buf), bph);
        if (elect_one()) mbar_arrive(tfull(buf));
        break;
      }
      const bool second = (entry & kP2Flag) != 0;
      const uint32_t idesc = second ? idesc2 : idesc1;
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        mbar_wait(tempty(buf), bph);
        tc_fence_after();
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
        if (++buf == kNumBuf) { buf = 0; bph ^= 1; }
      }
      if (++s == S) { s = 0; ph ^= 1; }
    }
  } else {
    // ===== epilogue warps: warpgroup g owns columns [32g, 32g+32) of every accumulator tile =====
    const int et = tid - 64;
    const int wg = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r_loc = quad * 32 + lane;
    const float scale = p.scale;
    // stage the Q block(s): canonical K-major SWIZZLE_128B layout
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

    float m_run[RB];
    float2 zacc[RB][4];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      m_run[rb] = -1e30f;
#pragma unroll
      for (int a = 0; a < 4; ++a) zacc[rb][a] = make_float2(0.f, 0.f);
    }
    const uint32_t tcol = tmem_base + ((uint32_t)(quad * 32) << 16) + wg * 32;
    // small windows (g*w < 128): accumulator rows beyond g*w are zero padding -- warps that own only padding rows
    // skip the load and the 32 exps per tile
    bool dead_rows[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) dead_rows[rb] = __all_sync(0xffffffffu, (rb * 128 + r_loc) / W >= p.groups);
    constexpr int kHeadsPerChunk = 32 / W;

    bool in_p1 = true;
    int n1_done = 0;     // pass-1 tiles this CTA has processed (uniform over the epilogue warps)
    unsigned aseq = 0;   // accumulator sequence number (RB per tile, 1 for the stop marker)
    for (int ts = 0;; ++ts) {
      mbar_wait(tfull(aseq % kNumBuf), (aseq / kNumBuf) & 1);
      tc_fence_after();
      const int entry = ring[ts % kRing];
      if (in_p1 && (entry < 0 || (entry & kP2Flag))) {
        // ---------------- end of this CTA's pass 1: publish, wait for the unit, fold ----------------
        in_p1 = false;
        if (n1_done > 0) {
#pragma unroll
          for (int rb = 0; rb < RB; ++rb) {
            const float2 z2 = __fadd2_rn(__fadd2_rn(zacc[rb][0], zacc[rb][1]), __fadd2_rn(zacc[rb][2], zacc[rb][3]));
            xstat[(rb * kNumWG + wg) * 128 + r_loc] = make_float2(m_run[rb], z2.x + z2.y);
          }
          if (et == 0) misc[0] = (int)atomicAdd(&ctl->nparts[unit], 1u);
          epi_sync();
          const int slot = misc[0];
          for (int r = et; r < kRows; r += kEpiThreads) {
            const int rb = r >> 7, rl = r & 127;
            float mn = -1e30f, z = 0.f;
#pragma unroll
            for (int g = 0; g < kNumWG; ++g) mn = fmaxf(mn, xstat[(rb * kNumWG + g) * 128 + rl].x);
#pragma unroll
            for (int g = 0; g < kNumWG; ++g) {
              const float2 a = xstat[(rb * kNumWG + g) * 128 + rl];
              z += a.y * fast_exp2((a.x - mn) * kLog2e);
            }
            p.stats[((int64_t)unit * p.slots_per_unit + slot) * kRows + r] = make_float2(mn, z);
          }
          __threadfence();
          epi_sync();
          if (et == 0) atomicAdd(&ctl->done[unit], (unsigned)n1_done);
        }
        if (entry < 0) break;   // this CTA got no pass-2 tile
        if (et == 0) {
          // every pass-1 tile of the unit has been claimed by a RUNNING CTA (else the producer could not have
          // moved on to pass 2), and those CTAs publish without waiting for anybody: this wait always ends
          for (unsigned spin = 0; ld_acquire_u32(&ctl->done[unit]) < (unsigned)T1; ++spin) {
            __nanosleep(40);
            if (spin > (1u << 24)) {
              printf("kvc scores_dyn: statistics of unit %d never completed (block %d)\n", unit, blockIdx.x);
              __trap();
            }
          }
          misc[1] = (int)ld_acquire_u32(&ctl->nparts[unit]);
        }
        epi_sync();
        {
          // fold every partial of the unit: thread = (row, part), part q folds partials q, q+kParts, ...
          const int nparts = misc[1];
          constexpr int kParts = kEpiThreads / kRows;
          const int r = et % kRows, q = et / kRows;
          const float2* st = p.stats + (int64_t)unit * p.slots_per_unit * kRows + r;
          float M = -1e30f, Z = 0.f;
          for (int c0 = q; c0 < nparts; c0 += 8 * kParts) {
            float2 sv[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) {
              sv[u] = make_float2(-1e30f, 0.f);
              if (c0 + u * kParts < nparts) {
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
          epi_sync();   // xstat is reused: everybody is past the publish reads
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
      }
      if (entry < 0) break;
      const int tile = entry & (kP2Flag - 1);
      if (!(entry & kP2Flag)) {
        // ---------------- pass 1: thread <-> query row, columns are keys ----------------
        const int key0 = tile * kTile + wg * 32;
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          const int buf = (aseq + rb) % kNumBuf;
          if (rb > 0) { mbar_wait(tfull(buf), ((aseq + rb) / kNumBuf) & 1); tc_fence_after(); }
          if (dead_rows[rb]) {
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
          if (key0 + 32 > nscore) {  // chunk touches the w x w causal block or the ragged end
            const int rg = rb * 128 + r_loc;
            const int last_key = min(p.L - 1, rg / W < p.groups ? nscore + rg % W : p.L - 1);
            p1_chunk<T, true, POLY1>(v, scale, m_run[rb], zacc[rb], key0, last_key);
          } else {
            p1_chunk<T, false, POLY1>(v, scale, m_run[rb], zacc[rb], 0, 0);
          }
        }
        ++n1_done;
      } else {
        // ---------------- pass 2: thread <-> key, columns are the stacked query rows ----------------
        const int key = tile * kTile + r_loc;
#pragma unroll
        for (int rb = 0; rb < RB; ++rb) {
          const int buf = (aseq + rb) % kNumBuf;
          if (rb > 0) { mbar_wait(tfull(buf), ((aseq + rb) / kNumBuf) & 1); tc_fence_after(); }
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
          float2 acc[kHeadsPerChunk];
#pragma unroll
          for (int h = 0; h < kHeadsPerChunk; ++h) acc[h] = make_float2(0.f, 0.f);
          p2_chunk<T, W, POLY2>(v, scale, crow + rb * 128 + wg * 32, acc);
          if (key < nscore) {
#pragma unroll
            for (int h = 0; h < kHeadsPerChunk; ++h) {
              const int head = (rb * 128 + wg * 32) / W + h;
              if (head < p.groups) {
                float s = acc[h].x + acc[h].y;
                if (p.reduce_mean) s = s / (float)p.w;
                p.out[((int64_t)b * p.Hq + kvh * p.groups + head) * nscore + key] = T::round1(s);
              }
            }
          }
        }
      }
      aseq += RB;
    }
  }

  tc_fence_before();
  __syncthreads();
  tl_end(p.tl);
  if (tid == 0) {
    // the last CTA to leave resets the control block for the next launch on this stream
    __threadfence();
    const unsigned old = atomicAdd(&ctl->exited, 1u);
    if (old == gridDim.x - 1) {
      for (int u = 0; u < p.units; ++u) { ctl->p1_next[u] = 0u; ctl->p2_next[u] = 0u; ctl->done[u] = 0u; ctl->nparts[u] = 0u; }
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
  int rb, units, grid, slots_per_unit;
  size_t stats_bytes;
};

static Plan make_plan(int B, int Hq, int Hkv, int L, int w) {
  Plan pl{};
  const int g = Hq / Hkv;
  pl.rb = (g * w + 127) / 128;
  pl.units = B * Hkv;
  const int tiles = (L + kTile - 1) / kTile;
  const int sms = sm_count();
  long long want = (long long)pl.units * tiles;      // never more CTAs than pass-1 tiles
  pl.grid = (int)(want < sms ? want : sms);
  if (pl.grid < pl.units) pl.grid = pl.units;
  pl.slots_per_unit = (pl.grid + pl.units - 1) / pl.units;   // a CTA publishes at most once
  pl.stats_bytes = ((size_t)pl.units * pl.slots_per_unit * pl.rb * 128 * sizeof(float2) + 255) & ~(size_t)255;
  return pl;
}

template <typename T, int RB, int W>
static int launch(const CUtensorMap& tmap, const Params& sp, const Plan& pl, cudaStream_t stream) {
  // Exponentials on the FMA pipe: 4 of 16 pairs in each pass (scripts/ubench/epi_mix.cu: 1086 -> 881 and
  // 1088 -> 939 cycles per tile of pure epilogue math); KVC_DYN_POLY=0 keeps every exponential on MUFU.
  static const bool poly = [] { const char* e = getenv("KVC_DYN_POLY"); return !(e && e[0] == '0'); }();
  auto kern = poly ? window_score_dyn_kernel<T, RB, W, 4, 4> : window_score_dyn_kernel<T, RB, W, 0, 0>;
  const int smem = Smem<RB>::kAlloc;
  KVC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
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

template <typename T, int RB>
static int dispatch_w(const CUtensorMap& tmap, const Params& sp, const Plan& pl, cudaStream_t stream) {
  switch (sp.w) {
    case 8: return launch<T, RB, 8>(tmap, sp, pl, stream);
    case 16: return launch<T, RB, 16>(tmap, sp, pl, stream);
    default: return launch<T, RB, 32>(tmap, sp, pl, stream);
  }
}

}  // namespace dyn

bool window_scores_tc_supported(int B, int Hq, int Hkv, int L, int d, int w);

// One CTA per SM serves one unit for its whole life: units must fit the SM count.
bool window_scores_dyn_supported(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_tc_supported(B, Hq, Hkv, L, d, w)) return false;
  if (w > 32) return false;  // a 64-row head spans two warpgroups' columns: the two-launch kernel covers it
  const int units = B * Hkv;
  return units <= sm_count() && units <= dyn::kMaxUnits;
}

size_t window_scores_dyn_workspace_bytes(int B, int Hq, int Hkv, int L, int d, int w) {
  if (!window_scores_dyn_supported(B, Hq, Hkv, L, d, w)) return 0;
  return dyn::make_plan(B, Hq, Hkv, L, w).stats_bytes + 256;
}

int window_scores_dyn(const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, const void* k, int64_t k_sb,
                      int64_t k_sh, int64_t k_sl, int B, int Hq, int Hkv, int L, int d, int w, int dtype, int reduce,
                      void* out_scores, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  using namespace dyn;
  KVC_REQUIRE(window_scores_dyn_supported(B, Hq, Hkv, L, d, w), KVC_ERR_UNSUPPORTED,
              "dynamic tcgen05 scoring kernel does not cover this shape (Hq/Hkv=%d, w=%d, d=%d, units=%d)", Hq / Hkv, w, d,
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
  sp.units = pl.units; sp.slots_per_unit = pl.slots_per_unit;
  sp.scale = 1.0f / sqrtf((float)d);
  sp.reduce_mean = reduce == KVC_REDUCE_MEAN;
  sp.stats = (float2*)workspace;
  sp.out = (uint16_t*)out_scores;
  sp.tl = KVC_TL_SLOT(0);
  if (dtype == KVC_BF16)
    return pl.rb == 1 ? dispatch_w<Bf16, 1>(tmap, sp, pl, stream) : dispatch_w<Bf16, 2>(tmap, sp, pl, stream);
  return pl.rb == 1 ? dispatch_w<Fp16, 1>(tmap, sp, pl, stream) : dispatch_w<Fp16, 2>(tmap, sp, pl, stream);
}

}  // namespace kvc
