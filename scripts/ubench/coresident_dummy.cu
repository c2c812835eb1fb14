// Feasibility probe: a sync-heavy, low-issue kernel with the footprint a co-resident select would have
// (clusters of 4 x 256 threads, <= 64 registers, 26 KB smem), built as a tiny shared library so that
// scripts/overlap_dummy.py can launch it on a side stream next to the scoring kernel.
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__global__ void __launch_bounds__(256) dummy_select(const uint16_t* __restrict__ scores, int n, int rounds, int* out) {
  extern __shared__ int sm[];
  int* hist = sm;  // 8 warps x 256
  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  uint32_t crank; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(crank));
  const uint16_t* row = scores + (size_t)blockIdx.y * n + crank * 8192 + tid * 32;
  uint4 q[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) q[i] = __ldg(reinterpret_cast<const uint4*>(row) + i);
  int acc = 0;
  for (int r = 0; r < rounds; ++r) {
    for (int i = lane; i < 256; i += 32) hist[wid * 256 + i] = 0;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t w4[4] = {q[i].x, q[i].y, q[i].z, q[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        atomicAdd(&hist[wid * 256 + ((w4[j] >> (r & 7)) & 255)], 1);
        atomicAdd(&hist[wid * 256 + ((w4[j] >> (16 + (r & 7))) & 255)], 1);
      }
    }
    __syncthreads();
    int s = 0;
    for (int w = 0; w < 8; ++w) s += hist[w * 256 + tid];
    acc += s;
    cluster_sync();
  }
  if (acc == -1) out[0] = acc;
}
extern "C" int dummy_launch(const void* scores, int rows, int n, int rounds, int* out, void* stream) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(4, rows); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = 26 * 1024; cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1]; attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 4; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  return (int)cudaLaunchKernelEx(&cfg, dummy_select, (const uint16_t*)scores, n, rounds, out);
}
