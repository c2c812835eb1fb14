// ThinK query-driven key-channel pruning of the COMPRESSED keys (reference key_pruner_query_driven,
// pyramidkv/pyramidkv_utils.py:13-26, called from SnapKVCluster.update_think :349-392):
//   queries_norm[c] = mean over the last 32 queries of q^2          (each square and the mean rounded to the model dtype)
//   keys_norm[c]    = mean over all compressed rows of k^2
//   key[c]          = queries_norm[c] * keys_norm[c]                (model dtype)
//   the int(d * ratio) channels with the smallest key are dropped from every row but the last `recent` ones;
//   kept channels stay in channel order; the boolean mask marks the kept channels.
// Two launches: (A) per (head, row chunk) partial column sums of the rounded squares, in a fixed order; (B) every CTA
// folds the partials of its head (deterministically, same bits in every CTA), ranks the channels and compacts its chunk
// of rows with 2-byte stores (the pruned tensor is d - k wide: no vector alignment to rely on).  The compressed cache is
// 16 MB at cfg-2: both launches are L2-resident right after the gather.
#include "common.cuh"

namespace kvc {

constexpr int kThinkQ = 32;          // the reference hard-codes the last 32 queries (:17)
constexpr int kThinkChunk = 64;      // rows per CTA

template <typename T>
__global__ void __launch_bounds__(256)
think_partial_kernel(const uint16_t* __restrict__ kc, int rows, int d, int nchunks, float* __restrict__ part) {
  const int bh = blockIdx.y, ch = blockIdx.x;
  const int r0 = ch * kThinkChunk, r1 = min(rows, r0 + kThinkChunk);
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float s = 0.f;
    const uint16_t* p = kc + ((int64_t)bh * rows + r0) * d + c;
    for (int r = r0; r < r1; ++r, p += d) {
      const float x = T::to_float(*p);
      s += T::to_float((uint16_t)T::round1(x * x));      // torch.pow(k, 2) in the model dtype
    }
    part[((int64_t)bh * nchunks + ch) * d + c] = s;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
think_prune_kernel(const uint16_t* __restrict__ kc, const uint16_t* __restrict__ q, int64_t q_sb, int64_t q_sh, int64_t q_sl,
                   int H, int L, int rows, int d, int recent, int kdrop, int nchunks, const float* __restrict__ part,
                   uint16_t* __restrict__ out, uint8_t* __restrict__ mask) {
  extern __shared__ float sm[];
  float* key = sm;                                   // [d]
  int* pos = reinterpret_cast<int*>(sm + d);         // [d]: position among the kept channels, -1 = dropped
  const int bh = blockIdx.y, b = bh / H, h = bh % H;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float ks = 0.f;
    for (int s = 0; s < nchunks; ++s) ks += part[((int64_t)bh * nchunks + s) * d + c];
    const float kn = T::to_float((uint16_t)T::round1(div_exact(ks, (float)rows, 1.0f / (float)rows)));
    float qs = 0.f;
    const int nq = min(kThinkQ, L);
    const uint16_t* qp = q + b * q_sb + h * q_sh + (int64_t)(L - nq) * q_sl + c;
    for (int r = 0; r < nq; ++r, qp += q_sl) {
      const float x = T::to_float(*qp);
      qs += T::to_float((uint16_t)T::round1(x * x));
    }
    const float qn = T::to_float((uint16_t)T::round1(div_exact(qs, (float)nq, 1.0f / (float)nq)));
    key[c] = T::to_float((uint16_t)T::round1(qn * kn));
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float mine = key[c];
    int rank = 0;
    for (int o = 0; o < d; ++o) rank += (key[o] < mine) || (key[o] == mine && o < c);
    pos[c] = rank < kdrop ? -1 : 0;                  // the kdrop smallest go (ties: lowest channel first)
  }
  __syncthreads();
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    if (pos[c] < 0) continue;
    int before = 0;
    for (int o = 0; o < c; ++o) before += pos[o] >= 0;
    key[c] = __int_as_float(before);                 // reuse: kept position (pos keeps the flag)
  }
  __syncthreads();
  if (blockIdx.x == 0)
    for (int c = threadIdx.x; c < d; c += blockDim.x) mask[(int64_t)bh * d + c] = pos[c] >= 0;
  const int prows = rows - recent, dk = d - kdrop;
  const int r0 = blockIdx.x * kThinkChunk, r1 = min(prows, r0 + kThinkChunk);
  for (int i = threadIdx.x; i < (r1 - r0) * d; i += blockDim.x) {
    const int r = r0 + i / d, c = i % d;
    if (pos[c] >= 0) out[((int64_t)bh * prows + r) * dk + __float_as_int(key[c])] = kc[((int64_t)bh * rows + r) * d + c];
  }
}

size_t think_prune_workspace_bytes(int BH, int rows, int d) {
  const int nchunks = (rows + kThinkChunk - 1) / kThinkChunk;
  return (size_t)BH * nchunks * d * sizeof(float) + 256;
}

int think_prune(const void* kc, const void* q, int64_t q_sb, int64_t q_sh, int64_t q_sl, int B, int H, int L, int rows, int d,
                int recent, int kdrop, int dtype, void* out_pruned, uint8_t* out_mask, void* workspace, size_t workspace_bytes,
                cudaStream_t stream) {
  KVC_REQUIRE(workspace_bytes >= think_prune_workspace_bytes(B * H, rows, d), KVC_ERR_WORKSPACE, "think_prune workspace: have %zu need %zu",
              workspace_bytes, think_prune_workspace_bytes(B * H, rows, d));
  const int nchunks = (rows + kThinkChunk - 1) / kThinkChunk;
  float* part = (float*)workspace;
  const dim3 grid(nchunks, B * H);
  const size_t smem = (size_t)d * 8;
  if (dtype == KVC_BF16) {
    think_partial_kernel<Bf16><<<grid, 256, 0, stream>>>((const uint16_t*)kc, rows, d, nchunks, part);
    if (int rc = launch_status("think_partial_kernel")) return rc;
    think_prune_kernel<Bf16><<<grid, 256, smem, stream>>>((const uint16_t*)kc, (const uint16_t*)q, q_sb, q_sh, q_sl, H, L, rows, d,
                                                         recent, kdrop, nchunks, part, (uint16_t*)out_pruned, out_mask);
  } else {
    think_partial_kernel<Fp16><<<grid, 256, 0, stream>>>((const uint16_t*)kc, rows, d, nchunks, part);
    if (int rc = launch_status("think_partial_kernel")) return rc;
    think_prune_kernel<Fp16><<<grid, 256, smem, stream>>>((const uint16_t*)kc, (const uint16_t*)q, q_sb, q_sh, q_sl, H, L, rows, d,
                                                         recent, kdrop, nchunks, part, (uint16_t*)out_pruned, out_mask);
  }
  return launch_status("think_prune_kernel");
}

}  // namespace kvc
