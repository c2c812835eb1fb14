// AdaKV head-adaptive budget allocation (pyramidkv_utils.py:706-719 + init_metadata :682-698).
//
// The reference sorts every head, ratio-normalises, runs one global topk(H*base) over all H*n
// normalised scores and counts winners per head.  Scores are 16-bit floats, so the global k-th
// largest value is found exactly from histograms of an order-preserving 16-bit key:
//   phase 1  per head   : a = round(v * ratio_h) -> key, 4096-bin histogram of key>>4
//   phase 2  one CTA    : sum histograms over heads, locate the bin holding the K-th largest
//   phase 3  per head   : 16-bin histogram of the low bits inside that bin + count above it
//   phase 4  one CTA    : exact threshold, per-head winner counts (ties go to the lowest head
//                         index first — torch.topk leaves tie order unspecified), floor + round,
//                         head_lens and the exclusive prefix cu_klen
// In the multi-GPU layout (one KV-head shard per GPU) the two histogram sums are the only data
// that has to cross GPUs (16 KiB + 64 B): see kvcache_factory_b200/sharding.py.
#include <cstddef>

#include "common.cuh"

namespace kvc {

constexpr int kAdaThreads = 1024;
constexpr int kAdaBins = 4096;

struct AdaHeader {
  int thr_bin, above_bin;   // phase 2
  int thr_key, count_gt;    // phase 4
};

struct AdaWorkspace {
  AdaHeader* hdr;
  uint16_t* akeys;   // [H, n]
  int* hist1;        // [H, 4096]
  int* hist2;        // [H, 16]
  int* above;        // [H]
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t adakv_allocate_workspace_bytes(int H, int n) {
  return align256(sizeof(AdaHeader)) + align256((size_t)H * n * 2) + align256((size_t)H * kAdaBins * 4) +
         align256((size_t)H * 16 * 4) + align256((size_t)H * 4);
}

static AdaWorkspace carve(void* ws, int H, int n) {
  char* p = (char*)ws;
  AdaWorkspace a;
  a.hdr = (AdaHeader*)p; p += align256(sizeof(AdaHeader));
  a.akeys = (uint16_t*)p; p += align256((size_t)H * n * 2);
  a.hist1 = (int*)p; p += align256((size_t)H * kAdaBins * 4);
  a.hist2 = (int*)p; p += align256((size_t)H * 16 * 4);
  a.above = (int*)p;
  return a;
}

// Rows are cut into chunks of kAdaChunk scores, one CTA each (grid = chunks x heads: 256 CTAs at 32 x 64K instead of
// one CTA per head); a CTA's shared-memory histogram is merged into the head's zero-initialised global one.
constexpr int kAdaChunk = 8 * kAdaThreads;

template <typename T>
__global__ void __launch_bounds__(kAdaThreads)
ada_phase1_kernel(const uint16_t* __restrict__ pooled, const float* __restrict__ sums, int n, int normalize,
                  uint16_t* __restrict__ akeys, int* __restrict__ hist1) {
  __shared__ int hist[kAdaBins];
  const int h = blockIdx.y, tid = threadIdx.x;
  for (int i = tid; i < kAdaBins; i += kAdaThreads) hist[i] = 0;
  __syncthreads();
  // ratio_weight = sum(top base) / sum(all), each rounded to the model dtype (:710)
  float ratio = 1.f;
  if (normalize) {
    const float top = T::to_float(T::round1(sums[2 * h])), all = T::to_float(T::round1(sums[2 * h + 1]));
    ratio = T::to_float(T::round1(top / all));
  }
  const uint16_t* row = pooled + (int64_t)h * n;
  const int j0 = blockIdx.x * kAdaChunk;
#pragma unroll
  for (int u = 0; u < kAdaChunk / kAdaThreads; ++u) {
    const int j = j0 + u * kAdaThreads + tid;
    if (j < n) {
      uint32_t bits = row[j];
      if (normalize) bits = T::round1(T::to_float((uint16_t)bits) * ratio);  // :711
      const uint32_t key = float16_bits_to_key(bits);
      akeys[(int64_t)h * n + j] = (uint16_t)key;
      atomicAdd(&hist[key >> 4], 1);
    }
  }
  __syncthreads();
  for (int i = tid; i < kAdaBins; i += kAdaThreads)
    if (hist[i]) atomicAdd(&hist1[(int64_t)h * kAdaBins + i], hist[i]);
}

// Sum per-head histograms and find the bin where the descending running count reaches K.
__global__ void __launch_bounds__(kAdaThreads)
ada_phase2_kernel(const int* __restrict__ hist1, int H, int K, AdaHeader* hdr) {
  __shared__ __align__(16) int tot[kAdaBins];
  __shared__ int warp_sum[32];
  const int tid = threadIdx.x;
  {  // thread t sums bins [4t, 4t+4) over the heads with 16-byte loads, 8 heads in flight
    int4 s = make_int4(0, 0, 0, 0);
    const int4* col = reinterpret_cast<const int4*>(hist1) + tid;
    int h = 0;
    for (; h + 8 <= H; h += 8) {
      int4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldg(col + (int64_t)(h + u) * (kAdaBins / 4));
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; h < H; ++h) {
      const int4 v = __ldg(col + (int64_t)h * (kAdaBins / 4));
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    *reinterpret_cast<int4*>(&tot[4 * tid]) = s;
  }
  __syncthreads();
  // thread t owns bins [4*(1023-t), +4): an ascending-thread scan walks down from the top bin
  const int b0 = 4 * (kAdaThreads - 1 - tid);
  const int cs[4] = {tot[b0 + 3], tot[b0 + 2], tot[b0 + 1], tot[b0]};
  const int mine = cs[0] + cs[1] + cs[2] + cs[3];
  const int lane = tid & 31, wid = tid >> 5;
  int inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sum[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    const int w = warp_sum[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_sum[lane] = winc - w;
  }
  __syncthreads();
  int cum = warp_sum[wid] + inc - mine;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    if (cum < K && cum + cs[q] >= K) { hdr->thr_bin = b0 + 3 - q; hdr->above_bin = cum; }
    cum += cs[q];
  }
}

__global__ void __launch_bounds__(kAdaThreads)
ada_phase3_kernel(const uint16_t* __restrict__ akeys, const int* __restrict__ hist1, int n, const AdaHeader* hdr,
                  int* __restrict__ hist2, int* __restrict__ above) {
  __shared__ int h2[16];
  __shared__ int s_above;
  const int h = blockIdx.y, tid = threadIdx.x;
  const int thr_bin = hdr->thr_bin;
  if (tid < 16) h2[tid] = 0;
  if (tid == 0) s_above = 0;
  __syncthreads();
  if (blockIdx.x == 0) {   // the head's count above the threshold bin, once
    int a = 0;
    for (int i = thr_bin + 1 + tid; i < kAdaBins; i += kAdaThreads) a += hist1[(int64_t)h * kAdaBins + i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((tid & 31) == 0 && a) atomicAdd(&s_above, a);
  }
  const int j0 = blockIdx.x * kAdaChunk;
#pragma unroll
  for (int u = 0; u < kAdaChunk / kAdaThreads; ++u) {
    const int j = j0 + u * kAdaThreads + tid;
    if (j < n) {
      const uint32_t key = akeys[(int64_t)h * n + j];
      if ((int)(key >> 4) == thr_bin) atomicAdd(&h2[key & 15], 1);
    }
  }
  __syncthreads();
  if (tid < 16 && h2[tid]) atomicAdd(&hist2[h * 16 + tid], h2[tid]);
  if (tid == 0 && blockIdx.x == 0) above[h] = s_above;
}

// One warp: lane <-> head (32 heads per round, carried totals across rounds), tiny tables.
__global__ void __launch_bounds__(32)
ada_phase4_kernel(const int* __restrict__ hist2, const int* __restrict__ above, int H, int n, int K,
                  float one_minus_floor, int floor_capacity, int w, AdaHeader* hdr,
                  int32_t* __restrict__ out_cap, int32_t* __restrict__ out_head_lens,
                  int32_t* __restrict__ out_cu_klen) {
  const int lane = threadIdx.x;
  // exact threshold: totals of the 16 low-bit bins over all heads (lanes 0..15 <-> bins)
  int tot = 0;
  if (lane < 16)
    for (int h = 0; h < H; ++h) tot += hist2[h * 16 + lane];
  int cum = hdr->above_bin, thr_low = 0;
  for (int t = 15; t >= 0; --t) {
    const int tt = __shfl_sync(0xffffffffu, tot, t);
    if (cum + tt >= K) { thr_low = t; break; }
    cum += tt;
  }
  if (lane == 0) { hdr->thr_key = (hdr->thr_bin << 4) | thr_low; hdr->count_gt = cum; }
  int remaining = K - cum;  // ties still to hand out, lowest head index first
  int run = 0;
  for (int h0 = 0; h0 < H; h0 += 32) {
    const int h = h0 + lane;
    const bool in = h < H;
    int cnt = 0, eq = 0;
    if (in) {
      cnt = above[h];
      for (int t = 15; t > thr_low; --t) cnt += hist2[h * 16 + t];
      eq = hist2[h * 16 + thr_low];
    }
    int eq_inc = eq;  // inclusive scan of the tie counts over the heads of this round
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, eq_inc, o);
      if (lane >= o) eq_inc += t;
    }
    const int before = eq_inc - eq;
    const int left = remaining - before;
    const int take = left <= 0 ? 0 : (eq < left ? eq : left);
    remaining -= min(remaining, __shfl_sync(0xffffffffu, eq_inc, 31));
    cnt += take;
    // torch.round(count * (1 - floor_ratio) + floor_capacity).int(): two fp32 roundings, half-to-even
    const float v = __fadd_rn(__fmul_rn((float)cnt, one_minus_floor), (float)floor_capacity);
    int cap = (int)rintf(v);
    if (cap > n) cap = n;  // sorted_idx[..., :cap] clamps at the row length
    if (cap < 0) cap = 0;
    const int len = in ? cap + w : 0;
    int len_inc = len;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, len_inc, o);
      if (lane >= o) len_inc += t;
    }
    if (in) {
      out_cap[h] = cap;
      out_head_lens[h] = cap + w;
      out_cu_klen[h] = run + len_inc - len;
    }
    run += __shfl_sync(0xffffffffu, len_inc, 31);
  }
  if (lane == 0) out_cu_klen[H] = run;
}

int adakv_allocate(const void* pooled, const float* sums, int H, int n, int dtype, int base_capacity,
                   float one_minus_floor, int floor_capacity, int normalize, int w, int32_t* out_cap,
                   int32_t* out_head_lens, int32_t* out_cu_klen, void* workspace, size_t workspace_bytes,
                   cudaStream_t stream) {
  KVC_REQUIRE(workspace_bytes >= adakv_allocate_workspace_bytes(H, n), KVC_ERR_WORKSPACE,
              "adakv_allocate workspace: have %zu need %zu", workspace_bytes, adakv_allocate_workspace_bytes(H, n));
  KVC_REQUIRE((int64_t)H * base_capacity <= (int64_t)H * n, KVC_ERR_ARG, "base capacity exceeds row length");
  AdaWorkspace a = carve(workspace, H, n);
  const int K = H * base_capacity;
  const dim3 grid((n + kAdaChunk - 1) / kAdaChunk, H);
  // hist1 and hist2 are adjacent in the workspace: one memset zeroes both accumulation targets
  KVC_CUDA_TRY(cudaMemsetAsync(a.hist1, 0, (size_t)((char*)a.above - (char*)a.hist1), stream));
  if (dtype == KVC_BF16)
    ada_phase1_kernel<Bf16><<<grid, kAdaThreads, 0, stream>>>((const uint16_t*)pooled, sums, n, normalize, a.akeys, a.hist1);
  else
    ada_phase1_kernel<Fp16><<<grid, kAdaThreads, 0, stream>>>((const uint16_t*)pooled, sums, n, normalize, a.akeys, a.hist1);
  if (int rc = launch_status("ada_phase1_kernel")) return rc;
  ada_phase2_kernel<<<1, kAdaThreads, 0, stream>>>(a.hist1, H, K, a.hdr);
  if (int rc = launch_status("ada_phase2_kernel")) return rc;
  ada_phase3_kernel<<<grid, kAdaThreads, 0, stream>>>(a.akeys, a.hist1, n, a.hdr, a.hist2, a.above);
  if (int rc = launch_status("ada_phase3_kernel")) return rc;
  ada_phase4_kernel<<<1, 32, 0, stream>>>(a.hist2, a.above, H, n, K, one_minus_floor, floor_capacity, w, a.hdr, out_cap,
                                          out_head_lens, out_cu_klen);
  return launch_status("ada_phase4_kernel");
}

}  // namespace kvc

// ---- the two counting phases on their own (multi-GPU: the caller all-reduces between them) --------------------
namespace kvc {

int adakv_hist_coarse(const void* pooled, const float* sums, int H, int n, int dtype, int normalize, uint16_t* akeys,
                      int32_t* hist1, cudaStream_t stream) {
  const dim3 grid((n + kAdaChunk - 1) / kAdaChunk, H);
  KVC_CUDA_TRY(cudaMemsetAsync(hist1, 0, (size_t)H * kAdaBins * sizeof(int32_t), stream));
  if (dtype == KVC_BF16)
    ada_phase1_kernel<Bf16><<<grid, kAdaThreads, 0, stream>>>((const uint16_t*)pooled, sums, n, normalize, akeys, hist1);
  else
    ada_phase1_kernel<Fp16><<<grid, kAdaThreads, 0, stream>>>((const uint16_t*)pooled, sums, n, normalize, akeys, hist1);
  return launch_status("ada_phase1_kernel");
}

int adakv_hist_fine(const uint16_t* akeys, const int32_t* hist1, int H, int n, const int32_t* thr_bin, int32_t* hist2,
                    int32_t* above, cudaStream_t stream) {
  static_assert(offsetof(AdaHeader, thr_bin) == 0, "phase 3 reads the threshold bin through the header's first word");
  const dim3 grid((n + kAdaChunk - 1) / kAdaChunk, H);
  KVC_CUDA_TRY(cudaMemsetAsync(hist2, 0, (size_t)H * 16 * sizeof(int32_t), stream));
  ada_phase3_kernel<<<grid, kAdaThreads, 0, stream>>>(akeys, hist1, n, reinterpret_cast<const AdaHeader*>(thr_bin), hist2, above);
  return launch_status("ada_phase3_kernel");
}

}  // namespace kvc

// ---- multi-GPU building block ------------------------------------------------------------------
// Full-resolution (65536-bin) histogram of each local head's ratio-normalised 16-bit keys.  Summed
// over heads and all-reduced over the ranks of a head-sharded layout it yields the exact global
// top-(H*base) threshold with ONE collective (256 KiB); see kvcache_factory_b200/sharding.py.
namespace kvc {

template <typename T>
__global__ void __launch_bounds__(256)
ada_key_histogram_kernel(const uint16_t* __restrict__ pooled, const float* __restrict__ sums, int n, int normalize,
                         int* __restrict__ hist) {
  const int h = blockIdx.y;
  float ratio = 1.f;
  if (normalize) {
    const float top = T::to_float(T::round1(sums[2 * h])), all = T::to_float(T::round1(sums[2 * h + 1]));
    ratio = T::to_float(T::round1(top / all));
  }
  const uint16_t* row = pooled + (int64_t)h * n;
  int* hh = hist + (int64_t)h * 65536;
  for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    uint32_t bits = row[j];
    if (normalize) bits = T::round1(T::to_float((uint16_t)bits) * ratio);
    atomicAdd(&hh[float16_bits_to_key(bits)], 1);
  }
}

int adakv_key_histogram(const void* pooled, const float* sums, int H, int n, int dtype, int normalize, int32_t* out_hist,
                        cudaStream_t stream) {
  KVC_CUDA_TRY(cudaMemsetAsync(out_hist, 0, (size_t)H * 65536 * sizeof(int32_t), stream));
  int gx = (n + 256 * 8 - 1) / (256 * 8);
  if (gx < 1) gx = 1;
  dim3 grid(gx, H);
  if (dtype == KVC_BF16)
    ada_key_histogram_kernel<Bf16><<<grid, 256, 0, stream>>>((const uint16_t*)pooled, sums, n, normalize, out_hist);
  else
    ada_key_histogram_kernel<Fp16><<<grid, 256, 0, stream>>>((const uint16_t*)pooled, sums, n, normalize, out_hist);
  return launch_status("ada_key_histogram_kernel");
}

}  // namespace kvc
