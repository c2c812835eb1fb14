// Pool (avg / max / none, stride 1, 'same') + per-row top-k select with torch.topk(sorted=True) order.
//
// Scores are 16-bit floats, so the pooled value maps onto a 16-bit order-preserving key and the
// k-th largest key is found EXACTLY with two histogram passes (4096 bins on the high 12 bits, 16
// bins on the low 4).  Elements above the threshold plus the lowest-index ties are compacted in
// index order and bitonic-sorted on a (key, ~index) composite: descending value, ascending index.
// One CTA per row (row = one (batch, query head)); the row is re-read from L2, never from HBM.
#include "common.cuh"

namespace kvc {

constexpr int kSelThreads = 1024;
constexpr int kSelWarps = kSelThreads / 32;
constexpr int kHistBins = 4096;
constexpr int kSortSmemMax = 8192;  // composites kept in shared memory up to this many (64 KiB)

struct SelShared {
  int hist[kHistBins];
  int hist2[16];
  int warp_tmp[kSelWarps];
  int warp_tmp2[kSelWarps];
  float fsum[kSelWarps];
  int thr_bin, above_bin, thr_key, count_gt;
  int run_gt, run_eq;
};

// Exclusive prefix sum over the block of one int per thread (also returns the block total).
__device__ __forceinline__ int block_exclusive_scan(int v, int* warp_tmp, int& total) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_tmp[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int w = warp_tmp[lane];
    int winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, winc, o);
      if (lane >= o) winc += t;
    }
    warp_tmp[lane] = winc - w;  // exclusive per warp
  }
  __syncthreads();
  const int base = warp_tmp[wid];
  // total = exclusive base of last warp + its inclusive sum
  __shared__ int s_total;
  if (threadIdx.x == kSelThreads - 1) s_total = base + inc;
  __syncthreads();
  total = s_total;
  return base + inc - v;
}

__device__ __forceinline__ float block_sum(float v, float* tmp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) tmp[wid] = v;
  __syncthreads();
  float r = 0.f;
  if (wid == 0) {
    r = tmp[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    if (lane == 0) tmp[0] = r;
  }
  __syncthreads();
  r = tmp[0];
  __syncthreads();
  return r;
}

// Pooled value (as a 16-bit pattern of the model dtype) at position j of a row.
// avg: fp32 sum of the window in ascending position order, zero padding, divided by kernel_size
// (count_include_pad), rounded once — ATen avg_pool2d semantics behind F.avg_pool1d
// (pyramidkv_utils.py:329).  max: -inf padding (pyramidkv_utils.py:331).
template <typename T>
__device__ __forceinline__ uint32_t pooled_bits(const uint16_t* __restrict__ row, int n, int j, int radius, int pool,
                                                float inv_dummy, int kernel_size) {
  if (pool == KVC_POOL_NONE) return row[j];
  if (pool == KVC_POOL_MAX) {
    uint32_t best_key = 0;
    for (int t = -radius; t <= radius; ++t) {
      const int p = j + t;
      if (p >= 0 && p < n) {
        const uint32_t key = float16_bits_to_key(row[p]);
        best_key = key > best_key ? key : best_key;
      }
    }
    return key_to_float16_bits(best_key);
  }
  float s = 0.f;
  for (int t = -radius; t <= radius; ++t) {
    const int p = j + t;
    if (p >= 0 && p < n) s += T::to_float(row[p]);
  }
  return T::round1(div_exact(s, (float)kernel_size, 1.0f / (float)kernel_size));
}

template <typename T>
__global__ void __launch_bounds__(kSelThreads)
pool_topk_kernel(const uint16_t* __restrict__ scores, int n, int kernel_size, int pool, int k_fixed,
                 const int32_t* __restrict__ k_per_row, int32_t* __restrict__ out_idx, int64_t idx_stride,
                 uint16_t* __restrict__ out_pooled, float* __restrict__ out_sums,
                 uint16_t* __restrict__ keys_scratch, unsigned long long* __restrict__ sort_scratch,
                 int64_t sort_scratch_stride, int sort_smem_cap) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SelShared& sh = *reinterpret_cast<SelShared*>(smem_raw);
  unsigned long long* sort_smem = reinterpret_cast<unsigned long long*>(smem_raw + sizeof(SelShared));

  const int row = blockIdx.x;
  const int tid = threadIdx.x;
  const uint16_t* srow = scores + (int64_t)row * n;
  uint16_t* keys = keys_scratch + (int64_t)row * n;
  int k = k_per_row ? min(k_per_row[row], k_fixed) : k_fixed;   // k_fixed is also the width of an out_idx row
  if (k > n) k = n;
  if (k < 0) k = 0;
  const int radius = kernel_size / 2;

  for (int i = tid; i < kHistBins; i += kSelThreads) sh.hist[i] = 0;
  if (tid < 16) sh.hist2[tid] = 0;
  if (tid == 0) { sh.run_gt = 0; sh.run_eq = 0; }
  __syncthreads();

  // ---- pass A: pool, key, 12-bit histogram ------------------------------------------------------
  float part_all = 0.f;
  for (int j = tid; j < n; j += kSelThreads) {
    const uint32_t bits = pooled_bits<T>(srow, n, j, radius, pool, 0.f, kernel_size);
    const uint32_t key = float16_bits_to_key(bits);
    keys[j] = (uint16_t)key;
    if (out_pooled) out_pooled[(int64_t)row * n + j] = (uint16_t)bits;
    atomicAdd(&sh.hist[key >> 4], 1);
    part_all += T::to_float((uint16_t)bits);
  }
  __syncthreads();
  float sum_all = 0.f;
  if (out_sums) sum_all = block_sum(part_all, sh.fsum);
  if (k == 0) {
    if (out_sums && tid == 0) { out_sums[2 * row] = 0.f; out_sums[2 * row + 1] = sum_all; }
    return;
  }

  // ---- threshold bin: first bin (from the top) where the running count reaches k ----------------
  {
    // thread t owns bins [4*(1023-t), +4) so that an ascending-thread scan walks DOWN from the top
    const int b0 = 4 * (kSelThreads - 1 - tid);
    const int c3 = sh.hist[b0 + 3], c2 = sh.hist[b0 + 2], c1 = sh.hist[b0 + 1], c0 = sh.hist[b0];
    int total;
    const int above = block_exclusive_scan(c3 + c2 + c1 + c0, sh.warp_tmp, total);
    int cum = above;
    const int cs[4] = {c3, c2, c1, c0};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      if (cum < k && cum + cs[q] >= k) { sh.thr_bin = b0 + 3 - q; sh.above_bin = cum; }
      cum += cs[q];
    }
  }
  __syncthreads();
  const int thr_bin = sh.thr_bin;

  // ---- pass B: low-4-bit histogram inside the threshold bin --------------------------------------
  for (int j = tid; j < n; j += kSelThreads) {
    const uint32_t key = keys[j];
    if ((int)(key >> 4) == thr_bin) atomicAdd(&sh.hist2[key & 15], 1);
  }
  __syncthreads();
  if (tid == 0) {
    int cum = sh.above_bin;
    for (int t = 15; t >= 0; --t) {
      if (cum + sh.hist2[t] >= k) { sh.thr_key = (thr_bin << 4) | t; sh.count_gt = cum; break; }
      cum += sh.hist2[t];
    }
  }
  __syncthreads();
  const uint32_t thr_key = (uint32_t)sh.thr_key;
  const int count_gt = sh.count_gt;
  const int need_eq = k - count_gt;

  // ---- pass C: ordered compaction into the sort buffer ------------------------------------------
  int kpad = 1;
  while (kpad < k) kpad <<= 1;
  // per-row k (AdaKV): a row with a small k sorts in shared memory even when the launch's largest k needs the
  // global scratch; the shared-memory sort area is sized by the host for min(kpad of the largest k, kSortSmemMax)
  unsigned long long* buf = (kpad <= sort_smem_cap) ? sort_smem : sort_scratch + (int64_t)row * sort_scratch_stride;
  constexpr int kItems = 4;
  for (int base = 0; base < n; base += kSelThreads * kItems) {
    const int j0 = base + tid * kItems;
    uint32_t kk[kItems];
    int gt = 0, eq = 0;
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      kk[u] = (j0 + u < n) ? keys[j0 + u] : 0u;
      const bool in = j0 + u < n;
      gt += (in && kk[u] > thr_key);
      eq += (in && kk[u] == thr_key);
    }
    int total;
    const int ex = block_exclusive_scan(gt | (eq << 16), sh.warp_tmp2, total);
    int gpos = sh.run_gt + (ex & 0xffff);
    int epos = sh.run_eq + (ex >> 16);
#pragma unroll
    for (int u = 0; u < kItems; ++u) {
      if (j0 + u < n) {
        const unsigned long long comp = ((unsigned long long)kk[u] << 32) | (unsigned long long)(0xffffffffu - (uint32_t)(j0 + u));
        if (kk[u] > thr_key) { buf[gpos++] = comp; }
        else if (kk[u] == thr_key) { if (epos < need_eq) buf[count_gt + epos] = comp; ++epos; }
      }
    }
    __syncthreads();
    if (tid == 0) { sh.run_gt += total & 0xffff; sh.run_eq += total >> 16; }
    __syncthreads();
  }
  for (int i = k + tid; i < kpad; i += kSelThreads) buf[i] = 0ull;
  __syncthreads();

  // ---- bitonic sort, descending ------------------------------------------------------------------
  for (int size = 2; size <= kpad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < (kpad >> 1); i += kSelThreads) {
        const int pos = 2 * i - (i & (stride - 1));
        const unsigned long long a = buf[pos], b = buf[pos + stride];
        const bool desc = (pos & size) == 0;
        if (desc ? (a < b) : (a > b)) { buf[pos] = b; buf[pos + stride] = a; }
      }
      __syncthreads();
    }
  }

  // ---- emit ------------------------------------------------------------------------------------
  float part_top = 0.f;
  for (int r = tid; r < k; r += kSelThreads) {
    const unsigned long long comp = buf[r];
    out_idx[(int64_t)row * idx_stride + r] = (int32_t)(0xffffffffu - (uint32_t)(comp & 0xffffffffull));
    part_top += T::to_float((uint16_t)key_to_float16_bits((uint32_t)(comp >> 32)));
  }
  if (out_sums) {
    const float sum_top = block_sum(part_top, sh.fsum);
    if (tid == 0) { out_sums[2 * row] = sum_top; out_sums[2 * row + 1] = sum_all; }
  }
}

int pool_topk_cluster(const void* scores, int rows, int n, int dtype, int kernel_size, int pool, int k,
                      const int32_t* k_per_row, int32_t* out_idx, int64_t idx_stride, void* out_pooled,
                      float* out_sums, cudaStream_t stream);  // select_cluster.cu; returns 1 when not applicable

static int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }

size_t pool_topk_workspace_bytes(int rows, int n, int k) {
  size_t keys = ((size_t)rows * n * 2 + 255) & ~(size_t)255;
  const int kpad = next_pow2(k > 1 ? k : 1);
  size_t sort = kpad > kSortSmemMax ? (size_t)rows * kpad * 8 : 0;
  return keys + sort;
}

int pool_topk(const void* scores, int rows, int n, int dtype, int kernel_size, int pool, int k,
              const int32_t* k_per_row, int32_t* out_idx, int64_t idx_stride, void* out_pooled, float* out_sums,
              void* workspace, size_t workspace_bytes, cudaStream_t stream) {
  KVC_REQUIRE(workspace_bytes >= pool_topk_workspace_bytes(rows, n, k), KVC_ERR_WORKSPACE,
              "pool_topk workspace: have %zu need %zu", workspace_bytes, pool_topk_workspace_bytes(rows, n, k));
  {
    const int rc = pool_topk_cluster(scores, rows, n, dtype, kernel_size, pool, k, k_per_row, out_idx, idx_stride,
                                     out_pooled, out_sums, stream);
    if (rc != 1) return rc;  // 1 = shape outside the cluster kernel: fall through to the one-CTA-per-row kernel
  }
  KVC_REQUIRE(out_idx != nullptr || k == 0, KVC_ERR_UNSUPPORTED,
              "sums-only select (out_idx == NULL with k > 0) is implemented by the cluster kernel only (rows=%d n=%d k=%d)",
              rows, n, k);
  const int kpad = next_pow2(k > 1 ? k : 1);
  const size_t keys_bytes = ((size_t)rows * n * 2 + 255) & ~(size_t)255;
  uint16_t* keys = (uint16_t*)workspace;
  unsigned long long* sort_scratch = kpad > kSortSmemMax ? (unsigned long long*)((char*)workspace + keys_bytes) : nullptr;
  const int sort_smem_cap = kpad > kSortSmemMax ? kSortSmemMax : kpad;
  const size_t smem = sizeof(SelShared) + (size_t)sort_smem_cap * 8;
  auto kern = dtype == KVC_BF16 ? pool_topk_kernel<Bf16> : pool_topk_kernel<Fp16>;
  if (dtype == KVC_BF16) { KVC_SET_MAX_SMEM(pool_topk_kernel<Bf16>, (int)smem); } else { KVC_SET_MAX_SMEM(pool_topk_kernel<Fp16>, (int)smem); }
  kern<<<rows, kSelThreads, smem, stream>>>((const uint16_t*)scores, n, kernel_size, pool, k, k_per_row, out_idx,
                                            idx_stride, (uint16_t*)out_pooled, out_sums, keys, sort_scratch, kpad,
                                            sort_smem_cap);
  return launch_status("pool_topk_kernel");
}

}  // namespace kvc
