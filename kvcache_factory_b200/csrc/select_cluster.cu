// Pool + top-k select, cluster version: one thread-block CLUSTER per score row, so a handful of rows
// (B*Hq = 32 at batch 1) still occupies most of the chip.
//
// CTA c of the C-CTA cluster owns the contiguous slice [c*SL, (c+1)*SL) of the row (SL = 1024*E).
//   1. load the slice + pooling halo into registers (16-byte vectors), pool, keep 16-bit keys in registers
//   2. k-th largest key in at most two counting rounds over a DATA-ADAPTIVE binning: the row's key range
//      [kmin, kmax] is split into <= 256 bins of 2^sh keys and the threshold bin is located.  If that bin's
//      keys fit the sort buffers next to the winners they simply go along (step 4 orders them and the first k
//      sorted entries are the answer); otherwise a second round resolves its 2^sh <= 256 keys.  Counting uses
//      warp-private shared-memory histograms (plain atomics: the adaptive bins spread the keys, so
//      same-address conflicts are rare); per-CTA histograms are read through distributed shared memory.
//   3. winners are compacted in index order (block scan; the bases of a CTA follow from the lower-ranked
//      CTAs' histograms, no further exchange) and written straight into CTA 0's shared memory with
//      st.shared::cluster; the other CTAs exit
//   4. CTA 0 orders them with a stable 2-pass LSD radix sort (O(k), ranks from match.any) and emits
//      indices by descending value, ascending position among equal values (torch.topk sorted=True,
//      deterministic on ties)
// Rows longer than 16 * 8192 scores or k > 8192 use the one-CTA-per-row kernel in select.cu.
#include <cstdlib>

#include "common.cuh"

namespace kvc {
namespace cl {

#ifdef KVC_TRACE
__device__ long long g_sel_trace[16];
__device__ long long g_sel_gt[8];
#define SEL_STAMP(slot) do { if (blockIdx.y == 0 && blockIdx.x == 0 && threadIdx.x == 0) g_sel_trace[slot] = clock64(); } while (0)
__device__ long long g_sel_tail[16];
#define TAIL_STAMP(slot) do { if (blockIdx.y == 0 && blockIdx.x == 3 && (threadIdx.x == blockDim.x - 1 || threadIdx.x == 0)) { asm volatile("" ::: "memory"); g_sel_tail[slot + (threadIdx.x == 0 ? 8 : 0)] = clock64(); } } while (0)
#else
#define TAIL_STAMP(slot) do { } while (0)
#define SEL_STAMP(slot) do { } while (0)
#endif

constexpr int kMaxK = 8192;   // 1024-thread variant; the 512-thread variant sorts at most kMaxK / 2 winners
constexpr int kHalo = 3;  // pooling kernels up to 7 taps

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t map_to_rank(const void* local, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local)), "r"(rank));
  return r;
}
__device__ __forceinline__ uint32_t ld_cluster_u32(uint32_t a) { uint32_t v; asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ float ld_cluster_f32(uint32_t a) { float v; asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void st_cluster_u32(uint32_t a, uint32_t v) { asm volatile("st.shared::cluster.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_f32(uint32_t a, float v) { asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ void st_cluster_u16(uint32_t a, uint16_t v) { asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(a), "h"(v) : "memory"); }

template <int kWarps>
struct SharedT {
  int whist[kWarps][256];   // warp-private counters (select rounds: bins; sort: digit offsets)
  int hist[2][256];         // this CTA's per-round histograms, read by the whole cluster
  int warp_tmp[32];
  int warp_tmp2[32];
  float fsum[32];
  int mm[2];                // min / max key of this CTA's slice
  float psum[2];            // partial sums (selected, all) of this CTA
  float cpsum[16][2];       // CTA 0 only: every CTA's partial sums, written remotely before the last cluster barrier
  int sel_digit, sel_above, sel_cnt; // result of the per-round threshold search (+ population of the threshold bin)
  int sel_labove, sel_lcnt;          // the same two counts restricted to the CTAs of lower rank (= lower positions)
  unsigned long long warp_tmp64[32];
};

// Exclusive block scan of one int per thread (kWarps warps, <= 32).
template <int kWarps>
__device__ __forceinline__ int block_scan_excl(int v, int* warp_tmp, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // warp_tmp free to overwrite
  if (lane == 31) warp_tmp[wid] = inc;
  __syncthreads();
  const int mine = lane < kWarps ? warp_tmp[lane] : 0;  // lane l holds warp l's total
  int pre = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += t;
  }
  total = __shfl_sync(0xffffffffu, pre, 31);
  const int base = __shfl_sync(0xffffffffu, pre - mine, wid);
  return base + inc - v;
}

// The same for a pair of counts packed into 64 bits (low word, high word; neither overflows 32 bits).
template <int kWarps>
__device__ __forceinline__ unsigned long long block_scan_excl64(unsigned long long v, unsigned long long* warp_tmp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();
  if (lane == 31) warp_tmp[wid] = inc;
  __syncthreads();
  const unsigned long long mine = lane < kWarps ? warp_tmp[lane] : 0ull;
  unsigned long long pre = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, pre, o);
    if (lane >= o) pre += t;
  }
  const unsigned long long base = __shfl_sync(0xffffffffu, pre - mine, wid);
  return base + inc - v;
}

template <int kWarps>
__device__ __forceinline__ float block_sum_f(float v, float* tmp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) tmp[wid] = v;
  __syncthreads();
  float r = lane < kWarps ? tmp[lane] : 0.f;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// Stable rank: add one occurrence of `digit` per lane to this warp's counters and return the
// counter value seen by this lane as if lanes were processed in lane order.
__device__ __forceinline__ int warp_rank_digit(int* wh, int digit, int lane) {
  const unsigned peers = __match_any_sync(0xffffffffu, digit);
  const int old = wh[digit];
  __syncwarp();
  const int rank = __popc(peers & ((1u << lane) - 1u));
  if (rank == 0) wh[digit] = old + __popc(peers);  // lowest peer lane updates the counter
  __syncwarp();
  return old + rank;
}

// Sum the cluster's per-CTA histograms of `round` and find the bin at which the descending running
// count reaches krem.  Every CTA computes the same answer.  Threads 0..255 <-> bins 255..0.
// The per-CTA histograms also tell every CTA how many keys above / inside the threshold bin lie in the CTAs of
// lower rank (l_above, l_in): the bases of the ordered compaction, without another exchange.
template <int kWarps>
__device__ __forceinline__ void cluster_threshold(SharedT<kWarps>& sh, int round, int krem, uint32_t csize, uint32_t crank,
                                                  int& digit, int& above, int& in_bin, int& l_above, int& l_in) {
  const int tid = threadIdx.x;
  const int d = 255 - tid;
  unsigned cnt = 0, low = 0;
  if (tid < 256)
    for (uint32_t c = 0; c < csize; ++c) {
      const unsigned v = ld_cluster_u32(map_to_rank(&sh.hist[round][d], c));
      cnt += v;
      if (c < crank) low += v;
    }
  const unsigned long long ex64 = block_scan_excl64<kWarps>((unsigned long long)cnt | ((unsigned long long)low << 32), sh.warp_tmp64);
  const int ex = (int)(unsigned)ex64;
  if (tid < 256 && ex < krem && ex + (int)cnt >= krem) {
    sh.sel_digit = d; sh.sel_above = ex; sh.sel_cnt = (int)cnt;
    sh.sel_labove = (int)(ex64 >> 32); sh.sel_lcnt = (int)low;
  }
  __syncthreads();
  digit = sh.sel_digit;
  above = sh.sel_above;
  in_bin = sh.sel_cnt;
  l_above = sh.sel_labove;
  l_in = sh.sel_lcnt;
}

// One counting round: bins of the participating keys -> warp-private atomics -> CTA histogram.
template <int E, int kWarps>
__device__ __forceinline__ void count_round(SharedT<kWarps>& sh, int round, const int (&bin)[E], int lane, int wid) {
  int* wh = sh.whist[wid];
#pragma unroll
  for (int i = lane; i < 256; i += 32) wh[i] = 0;
  __syncwarp();
#pragma unroll
  for (int i = 0; i < E; ++i)
    if (bin[i] >= 0) atomicAdd(&wh[bin[i]], 1);
  __syncthreads();
  if (threadIdx.x < 256) {
    int s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sh.whist[w][threadIdx.x];
    sh.hist[round][threadIdx.x] = s;
  }
}

template <typename T, int E, int POOL, int kThreads>
__global__ void __launch_bounds__(kThreads, kThreads == 512 ? 2 : 1)
topk_cluster_kernel(const uint16_t* __restrict__ scores, int n, int kernel_size, int k_fixed,
                    const int32_t* __restrict__ k_per_row, int32_t* __restrict__ out_idx, int64_t idx_stride,
                    uint16_t* __restrict__ out_pooled, float* __restrict__ out_sums, unsigned long long* tl) {
  // programmatic dependent launch (KVC_PDL_TAIL): the grid may be scheduled while the kernel that produces the scores
  // is still draining; nothing is read before this point
  asm volatile("griddepcontrol.wait;" ::: "memory");
  tl_start(tl);
  constexpr int SL = kThreads * E;          // slice length
  constexpr int kWarps = kThreads / 32;
  constexpr int kMaxKT = kThreads == 1024 ? kMaxK : kMaxK / 2;   // winners the sort buffers of this variant hold
  using Shared = SharedT<kWarps>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Shared& sh = *reinterpret_cast<Shared*>(smem_raw);
  unsigned char* dyn = smem_raw + ((sizeof(Shared) + 15) & ~15);
  // phase 3-4 (CTA 0 only, written remotely): keys/indices ping-pong
  uint16_t* skey[2] = {reinterpret_cast<uint16_t*>(dyn), reinterpret_cast<uint16_t*>(dyn) + kMaxKT};
  uint32_t* sidx[2] = {reinterpret_cast<uint32_t*>(dyn + 4 * kMaxKT), reinterpret_cast<uint32_t*>(dyn + 4 * kMaxKT) + kMaxKT};

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const uint32_t crank = cluster_rank(), csize = cluster_size();
  const int row = blockIdx.y;
  const uint16_t* srow = scores + (int64_t)row * n;
  int k = k_per_row ? min(k_per_row[row], k_fixed) : k_fixed;   // k_fixed bounds the sort buffers and an out_idx row
  k = min(max(k, 0), n);
  const int radius = POOL == KVC_POOL_NONE ? 0 : kernel_size / 2;
  const int s0 = (int)crank * SL;
  const float ksf = (float)kernel_size, krcp = 1.0f / ksf;

  SEL_STAMP(0);
#ifdef KVC_TRACE
  if (blockIdx.y == 0 && threadIdx.x == 0 && blockIdx.x < 4) {
    unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    g_sel_gt[blockIdx.x] = (long long)g;
  }
#endif
  // ---- 1. load this thread's E scores (+ pooling halo) straight into registers, pool -------------
  // Own elements: 8-element (16-byte) vector loads when the row allows it -- a warp then reads one
  // contiguous span; the 2*kHalo halo elements are the neighbours' own elements (L1 hits).
  const int j0 = s0 + tid * E;
  uint32_t keys[E];
  float part_all = 0.f;
  {
    uint16_t raw[E + 2 * kHalo] = {};
    const bool vec =(n & 7) == 0 && (E & 7) == 0 && (reinterpret_cast<uintptr_t>(srow) & 15) == 0;
    if (vec) {
#pragma unroll
      for (int v8 = 0; v8 < E / 8; ++v8) {
        uint4 q = make_uint4(0u, 0u, 0u, 0u);
        if (j0 + v8 * 8 < n) q = __ldg(reinterpret_cast<const uint4*>(srow + j0 + v8 * 8));
        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          raw[kHalo + v8 * 8 + 2 * i] = (uint16_t)(w4[i] & 0xffffu);
          raw[kHalo + v8 * 8 + 2 * i + 1] = (uint16_t)(w4[i] >> 16);
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < E; ++i) raw[kHalo + i] = j0 + i < n ? __ldg(srow + j0 + i) : (uint16_t)0;
    }
    TAIL_STAMP(0);
    if (POOL != KVC_POOL_NONE) {
#pragma unroll
      for (int i = 0; i < kHalo; ++i) {
        const int jl = j0 - kHalo + i, jr = j0 + E + i;
        raw[i] = (jl >= 0 && jl < n) ? __ldg(srow + jl) : (uint16_t)0;
        raw[kHalo + E + i] = jr < n ? __ldg(srow + jr) : (uint16_t)0;
      }
    }
    // slots outside the row hold zero bits (= +0.0): the zero padding of avg_pool1d needs no test; max pooling pads
    // with the lowest key instead
    TAIL_STAMP(1);
    if (POOL == KVC_POOL_AVG) {
      float f[E + 2 * kHalo];
#pragma unroll
      for (int i = 0; i < E + 2 * kHalo; ++i) f[i] = T::to_float(raw[i]);
#pragma unroll
      for (int i = 0; i < E; ++i) {
        float s = 0.f;  // ascending position, divisor = kernel_size (ATen avg_pool)
        if (radius >= 3) s += f[i];
        if (radius >= 2) s += f[i + 1];
        if (radius >= 1) s += f[i + 2];
        s += f[i + 3];
        if (radius >= 1) s += f[i + 4];
        if (radius >= 2) s += f[i + 5];
        if (radius >= 3) s += f[i + 6];
        keys[i] = float16_bits_to_key(T::round1(div_exact(s, ksf, krcp)));
      }
    } else {
      uint32_t kk[E + 2 * kHalo];
#pragma unroll
      for (int i = 0; i < E + 2 * kHalo; ++i) {
        const int j = j0 - kHalo + i;
        kk[i] = (j >= 0 && j < n) ? float16_bits_to_key(raw[i]) : 0u;
      }
#pragma unroll
      for (int i = 0; i < E; ++i) {
        uint32_t best = kk[i + 3];
        if (POOL == KVC_POOL_MAX) {
          if (radius >= 1) best = max(best, max(kk[i + 2], kk[i + 4]));
          if (radius >= 2) best = max(best, max(kk[i + 1], kk[i + 5]));
          if (radius >= 3) best = max(best, max(kk[i], kk[i + 6]));
        }
        keys[i] = best;
      }
    }
    TAIL_STAMP(2);
    if (j0 + E > n) {   // ragged end of the row (warp-uniform except in one warp)
#pragma unroll
      for (int i = 0; i < E; ++i)
        if (j0 + i >= n) keys[i] = 0u;
    }
    if (out_pooled) {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if (j0 + i < n) out_pooled[(int64_t)row * n + j0 + i] = (uint16_t)key_to_float16_bits(keys[i]);
    }
    if (out_sums) {
#pragma unroll
      for (int i = 0; i < E; ++i)
        if (j0 + i < n) part_all += T::to_float((uint16_t)key_to_float16_bits(keys[i]));
    }
  }
    TAIL_STAMP(3);
  SEL_STAMP(1);
  const int nvalid = max(0, min(E, n - j0));
  const float sum_all_cta = out_sums ? block_sum_f<kWarps>(part_all, sh.fsum) : 0.f;
  if (k == 0) {  // uniform over the cluster; nothing to select (AdaKV asks only for pooled scores / sums)
    if (out_sums) {
      if (tid == 0) sh.psum[1] = sum_all_cta;
      cluster_sync();
      if (crank == 0 && tid == 0) {
        float a = 0.f;
        for (uint32_t c = 0; c < csize; ++c) a += ld_cluster_f32(map_to_rank(&sh.psum[1], c));
        out_sums[2 * row] = 0.f;
        out_sums[2 * row + 1] = a;
      }
      cluster_sync();
    }
    return;
  }

  SEL_STAMP(2);
  // ---- 2a. key range of the row ---------------------------------------------------------------------
  int kmin = 0xffff, kmax = 0;
#pragma unroll
  for (int i = 0; i < E; ++i)
    if (i < nvalid) { kmin = min(kmin, (int)keys[i]); kmax = max(kmax, (int)keys[i]); }
    TAIL_STAMP(4);
  kmin = __reduce_min_sync(0xffffffffu, kmin);
  kmax = __reduce_max_sync(0xffffffffu, kmax);
  if (lane == 0) { sh.warp_tmp[wid] = kmin; sh.warp_tmp2[wid] = kmax; }
  __syncthreads();
  if (wid == 0) {
    const int a = __reduce_min_sync(0xffffffffu, lane < kWarps ? sh.warp_tmp[lane] : 0xffff);
    const int b2 = __reduce_max_sync(0xffffffffu, lane < kWarps ? sh.warp_tmp2[lane] : 0);
    if (lane == 0) { sh.mm[0] = a; sh.mm[1] = b2; }
  }
  SEL_STAMP(9);
#ifdef KVC_TRACE
  if (blockIdx.y == 0 && threadIdx.x == 0 && blockIdx.x < 4) {
    unsigned long long g; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g));
    g_sel_gt[4 + blockIdx.x] = (long long)g;
  }
#endif
  cluster_sync();
  SEL_STAMP(10);
  kmin = 0xffff; kmax = 0;
  for (uint32_t c = 0; c < csize; ++c) {
    kmin = min(kmin, (int)ld_cluster_u32(map_to_rank(&sh.mm[0], c)));
    kmax = max(kmax, (int)ld_cluster_u32(map_to_rank(&sh.mm[1], c)));
  }
  const int range = kmax - kmin;
  const int sh1 = range < 256 ? 0 : (32 - __clz(range)) - 8;  // bins of 2^sh1 keys, <= 256 bins

  SEL_STAMP(3);
  // ---- 2b. counting rounds -----------------------------------------------------------------------------
  // After the rounds: keys >= thr_hi are winners (count_gt of them), keys in [thr_key, thr_hi) are CANDIDATES of
  // which the first need_eq in (descending key, ascending position) order complete the k.  An exact threshold has
  // thr_hi = thr_key + 1.  When the whole threshold bin of round one fits the sort buffers next to the winners, round
  // two is skipped: the bin's keys all go to the sort, which orders them anyway, and the first k sorted entries are
  // the answer (one counting round, one cluster barrier and one threshold search less on the latency chain).
  int krem = k, thr_key, thr_hi, nsort = k;
  int gbase, ebase;   // winners / candidates held by the CTAs of lower rank
  {
    int bin[E];
#pragma unroll
    for (int i = 0; i < E; ++i) bin[i] = i < nvalid ? ((int)keys[i] - kmin) >> sh1 : -1;
    count_round<E, kWarps>(sh, 0, bin, lane, wid);
    SEL_STAMP(11);
    cluster_sync();  // every CTA's hist[0] is complete and visible
    SEL_STAMP(12);
    int b1, above, in_bin;
    cluster_threshold<kWarps>(sh, 0, krem, csize, crank, b1, above, in_bin, gbase, ebase);
    krem -= above;
    thr_key = kmin + (b1 << sh1);
    thr_hi = thr_key + (1 << sh1);
    if (sh1 > 0) {
      if (out_idx != nullptr && out_sums == nullptr && above + in_bin <= kMaxKT) {
        krem = in_bin;              // take the whole bin along
        nsort = above + in_bin;
      } else {
#pragma unroll
        for (int i = 0; i < E; ++i) bin[i] = (bin[i] == b1) ? ((int)keys[i] - thr_key) : -1;
        count_round<E, kWarps>(sh, 1, bin, lane, wid);
        cluster_sync();
        int b2v, above2, in2, l_above2;
        cluster_threshold<kWarps>(sh, 1, krem, csize, crank, b2v, above2, in2, l_above2, ebase);
        gbase += l_above2;
        krem -= above2;
        thr_key += b2v;
        thr_hi = thr_key + 1;
      }
    }
  }
  SEL_STAMP(4);
  const int need_eq = krem, count_gt = nsort - krem;

  if (out_idx == nullptr) {
    // Sums only (AdaKV's ratio normalisation, pyramidkv_utils.py:706-711): the sum of the k largest
    // pooled values follows from the threshold alone -- no compaction, no sort, no index output.
    float part_top = 0.f;
#pragma unroll
    for (int i = 0; i < E; ++i)
      if (i < nvalid && (int)keys[i] > thr_key) part_top += T::to_float((uint16_t)key_to_float16_bits(keys[i]));
    const float st = block_sum_f<kWarps>(part_top, sh.fsum);
    if (tid == 0) { sh.psum[0] = st; sh.psum[1] = sum_all_cta; }
    cluster_sync();
    if (crank == 0 && tid == 0 && out_sums) {
      float a = 0.f, bsum = 0.f;
      for (uint32_t c = 0; c < csize; ++c) {
        a += ld_cluster_f32(map_to_rank(&sh.psum[0], c));
        bsum += ld_cluster_f32(map_to_rank(&sh.psum[1], c));
      }
      const float tv = T::to_float((uint16_t)key_to_float16_bits((uint32_t)thr_key));
      for (int e = 0; e < need_eq; ++e) a += tv;   // the ties taken at the threshold all carry its value
      out_sums[2 * row] = a;
      out_sums[2 * row + 1] = bsum;
    }
    cluster_sync();
    tl_end(tl);
    return;
  }

  // ---- 3. ordered compaction into CTA 0 -----------------------------------------------------------
  int gt = 0, eq = 0;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    gt += (i < nvalid && (int)keys[i] >= thr_hi);
    eq += (i < nvalid && (int)keys[i] >= thr_key && (int)keys[i] < thr_hi);
  }
  int total;
  const int ex = block_scan_excl<kWarps>(gt | (eq << 16), sh.warp_tmp, total);
  SEL_STAMP(5);
  int gpos = gbase + (ex & 0xffff), epos = ebase + (int)((uint32_t)ex >> 16);
  const uint32_t rkey = map_to_rank(skey[0], 0), ridx = map_to_rank(sidx[0], 0);
  epos += count_gt;
  const int eend = count_gt + need_eq;
#pragma unroll
  for (int i = 0; i < E; ++i) {
    const int ky = (int)keys[i];
    const bool is_gt = i < nvalid && ky >= thr_hi;
    const bool is_eq = i < nvalid && !is_gt && ky >= thr_key;
    const int pos = is_gt ? gpos : epos;
    if (is_gt || (is_eq && epos < eend)) {
      st_cluster_u16(rkey + 2u * pos, (uint16_t)ky);
      st_cluster_u32(ridx + 4u * pos, (uint32_t)(j0 + i));
    }
    gpos += is_gt;
    epos += is_eq;
  }
  if (out_sums) {   // exact threshold here (the bin shortcut is not taken when sums are requested)
    float part_top = 0.f;
    int e2 = ebase + (int)((uint32_t)ex >> 16);
#pragma unroll
    for (int i = 0; i < E; ++i) {
      const int ky = (int)keys[i];
      const bool is_gt = i < nvalid && ky >= thr_hi;
      const bool is_eq = i < nvalid && ky == thr_key;
      if (is_gt || (is_eq && e2 < need_eq)) part_top += T::to_float((uint16_t)key_to_float16_bits(keys[i]));
      e2 += is_eq;
    }
    const float st = block_sum_f<kWarps>(part_top, sh.fsum);
    if (tid == 0) {
      st_cluster_f32(map_to_rank(&sh.cpsum[crank][0], 0), st);
      st_cluster_f32(map_to_rank(&sh.cpsum[crank][1], 0), sum_all_cta);
    }
  }
  SEL_STAMP(13);
  cluster_sync();  // all winners (and partial sums) have landed in CTA 0
  if (crank != 0) { tl_end(tl); return; }   // nobody reads this CTA's shared memory any more: free the SM for the next kernel

  SEL_STAMP(6);
  // ---- 4. CTA 0: stable LSD radix sort by descending key, emit -------------------------------------
  {
    if (out_sums && tid == 0) {
      float a = 0.f, bsum = 0.f;
      for (uint32_t c = 0; c < csize; ++c) { a += sh.cpsum[c][0]; bsum += sh.cpsum[c][1]; }
      out_sums[2 * row] = a;
      out_sums[2 * row + 1] = bsum;
    }
    int* wh = sh.whist[wid];
    const int V = (nsort + kThreads - 1) / kThreads;      // slots per thread (<= 8)
    const int per_warp = V * 32;                          // warp w owns elements [w*per_warp, (w+1)*per_warp)
    int cur = 0;
    // LSD radix sort on rk = kmax - key (ascending rk == descending key); the winners' keys usually
    // span fewer than 256 values, then one pass sorts them completely
    const int npass = (kmax - thr_key) < 256 ? 1 : 2;
    for (int pass = 0; pass < npass; ++pass) {
#pragma unroll
      for (int i = lane; i < 256; i += 32) wh[i] = 0;
      __syncwarp();
      // per-warp digit totals (elements of a warp's run are visited in order, so ranks are stable)
      for (int v = 0; v < V; ++v) {
        const int o = wid * per_warp + v * 32 + lane;
        if (o < nsort) atomicAdd(&wh[((kmax - (int)skey[cur][o]) >> (8 * pass)) & 255], 1);
      }
      __syncthreads();
      // digit bases: thread t < 256 <-> digit t; exclusive scan over digits of the per-digit totals,
      // then the per-warp counters become start offsets: whist[w][d] = base[d] + sum_{w'<w} count[w'][d]
      int tot = 0;
      if (tid < 256) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) tot += sh.whist[w][tid];
      }
      int total2;
      int dbase = block_scan_excl<kWarps>(tot, sh.warp_tmp, total2);
      if (tid < 256) {
#pragma unroll
        for (int w = 0; w < kWarps; ++w) { const int c = sh.whist[w][tid]; sh.whist[w][tid] = dbase; dbase += c; }
      }
      __syncthreads();
      // scatter in (warp, slot, lane) order with stable in-warp ranks
      for (int v = 0; v < V; ++v) {
        const int o = wid * per_warp + v * 32 + lane;
        const bool in = o < nsort;
        const uint16_t key = in ? skey[cur][o] : (uint16_t)0;
        const uint32_t idx = in ? sidx[cur][o] : 0u;
        // lanes past the end use the last digit: they rank after every real element of the warp's run
        const int dgt = in ? ((kmax - (int)key) >> (8 * pass)) & 255 : 255;
        const int dst = warp_rank_digit(wh, dgt, lane);
        if (in) { skey[cur ^ 1][dst] = key; sidx[cur ^ 1][dst] = idx; }
      }
      __syncthreads();
      cur ^= 1;
    }
    SEL_STAMP(7);
    for (int r = tid; r < k; r += kThreads) out_idx[(int64_t)row * idx_stride + r] = (int32_t)sidx[cur][r];
  }
  SEL_STAMP(8);
  tl_end(tl);
}

template <typename T, int E, int POOL, int kThreads = 1024>
static int launch(const void* scores, int rows, int n, int C, int kernel_size, int k, const int32_t* k_per_row,
                  int32_t* out_idx, int64_t idx_stride, void* out_pooled, float* out_sums, cudaStream_t stream) {
  auto kern = topk_cluster_kernel<T, E, POOL, kThreads>;
  using Shared = SharedT<kThreads / 32>;
  const size_t stage_bytes = (size_t)(kThreads * E + 2 * kHalo + 2 * (kThreads + 2)) * 2;
  const size_t sort_bytes = (size_t)(kThreads == 1024 ? kMaxK : kMaxK / 2) * 12;
  const size_t smem = ((sizeof(Shared) + 15) & ~(size_t)15) + (stage_bytes > sort_bytes ? stage_bytes : sort_bytes);
  KVC_SET_MAX_SMEM(kern, (int)smem);
  if (C > 8) KVC_CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(C, rows);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cudaLaunchAttribute attr2[2];
  static const bool pdl = [] { const char* e = getenv("KVC_PDL_TAIL"); return e && atoi(e) == 1; }();
  attr2[0] = attr[0];
  attr2[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr2[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = pdl ? attr2 : attr;
  cfg.numAttrs = pdl ? 2 : 1;
  const uint16_t* sc = (const uint16_t*)scores;
  uint16_t* op = (uint16_t*)out_pooled;
  unsigned long long* tl = KVC_TL_SLOT(1);
  KVC_CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, sc, n, kernel_size, k, k_per_row, out_idx, idx_stride, op, out_sums, tl));
  return launch_status("topk_cluster_kernel");
}

template <typename T, int POOL>
static int dispatch_e(const void* scores, int rows, int n, int kernel_size, int k, const int32_t* k_per_row,
                      int32_t* out_idx, int64_t idx_stride, void* out_pooled, float* out_sums, cudaStream_t stream) {
  // slice = 1024*E scores per CTA
  if (n <= 4096) return launch<T, 4, POOL>(scores, rows, n, 1, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
  if (n <= 16384) return launch<T, 4, POOL>(scores, rows, n, (n + 4095) / 4096, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
  // 512-thread variant: 64 registers and ~67 KB of shared memory, so TWO CTAs fit on an SM and the selects of
  // consecutive layers (issued on rotating side streams by the host mirror) overlap their latency chains:
  // 10.4 us per layer for two overlapped selects against 15.8 us back to back with the 1024-thread variant,
  // 16.5 us when alone.  With 3 CTAs per SM (40 registers) the spills cost more than the overlap gains; 8-CTA clusters of
  // 512 x 8 (two CTAs of ONE launch per SM) are slower as well: 17.1 vs 13.5 us alone, 12.4 vs 9.2 us in pairs (round 2).
  // KVC_SELECT_THREADS=1024 forces the one-CTA-per-SM variant.
  static const bool half = [] { const char* e = getenv("KVC_SELECT_THREADS"); return !(e && atoi(e) == 1024); }();
  // (for longer rows the 8- and 16-CTA clusters of this variant were measured slower than 1024 x 16: 77 vs 69 us
  // per cfg-5 shard layer)
  if (half && n > 16384 && n <= 32768 && k <= kMaxK / 2)
    return launch<T, 16, POOL, 512>(scores, rows, n, (n + 8191) / 8192, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
  // one CTA occupies a whole SM (1 024 threads): prefer the slice length that keeps rows * C within one wave
  if (n <= 32768 || (n <= 65536 && rows * ((n + 8191) / 8192) <= sm_count()))
    return launch<T, 8, POOL>(scores, rows, n, (n + 8191) / 8192, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
  return launch<T, 16, POOL>(scores, rows, n, (n + 16383) / 16384, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
}

template <typename T>
static int dispatch_pool(int pool, const void* scores, int rows, int n, int kernel_size, int k, const int32_t* k_per_row,
                         int32_t* out_idx, int64_t idx_stride, void* out_pooled, float* out_sums, cudaStream_t stream) {
  if (pool == KVC_POOL_AVG) return dispatch_e<T, KVC_POOL_AVG>(scores, rows, n, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
  if (pool == KVC_POOL_MAX) return dispatch_e<T, KVC_POOL_MAX>(scores, rows, n, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
  return dispatch_e<T, KVC_POOL_NONE>(scores, rows, n, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
}

}  // namespace cl

#ifdef KVC_TRACE
extern "C" int kvc_debug_sel_trace(long long* out) { return (int)cudaMemcpyFromSymbol(out, cl::g_sel_trace, sizeof(long long) * 16); }
extern "C" int kvc_debug_sel_tail(long long* out) { return (int)cudaMemcpyFromSymbol(out, cl::g_sel_tail, sizeof(long long) * 16); }
extern "C" int kvc_debug_sel_gt(long long* out) { return (int)cudaMemcpyFromSymbol(out, cl::g_sel_gt, sizeof(long long) * 8); }
#endif

// Returns 1 when the shape is outside what the cluster kernel covers (caller falls back).
int pool_topk_cluster(const void* scores, int rows, int n, int dtype, int kernel_size, int pool, int k,
                      const int32_t* k_per_row, int32_t* out_idx, int64_t idx_stride, void* out_pooled, float* out_sums,
                      cudaStream_t stream) {
  if (k > cl::kMaxK || n > 8 * 16384 || (pool != KVC_POOL_NONE && kernel_size > 2 * cl::kHalo + 1) || rows > 65535)
    return 1;
  return dtype == KVC_BF16
             ? cl::dispatch_pool<Bf16>(pool, scores, rows, n, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream)
             : cl::dispatch_pool<Fp16>(pool, scores, rows, n, kernel_size, k, k_per_row, out_idx, idx_stride, out_pooled, out_sums, stream);
}

}  // namespace kvc
