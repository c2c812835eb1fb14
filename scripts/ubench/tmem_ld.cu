// TMEM read throughput: NW warps each read a [32 lanes x 128 columns] fp32 block (4 x tcgen05.ld.32x32b.x32) per "tile".
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>
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
__global__ void k(unsigned* out, long long* cyc, int iters) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tptr)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tptr + ((uint32_t)((warp & 3) * 32) << 16);
  unsigned acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[4][32];
    const uint32_t buf = (it & 3) * 128;
#pragma unroll
    for (int cc = 0; cc < 4; ++cc) tmem_ld32(base + buf + cc * 32, v[cc]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int cc = 0; cc < 4; ++cc)
#pragma unroll
      for (int i = 0; i < 32; ++i) acc ^= v[cc][i];
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tptr), "r"(512u) : "memory");
}
int main() {
  unsigned* d; long long* c; cudaMalloc(&d, 148 * 512 * 4); cudaMalloc(&c, 8);
  for (int nw : {4, 8, 12, 16}) {
    const int iters = 2000;
    k<<<148, nw * 32>>>(d, c, iters); cudaDeviceSynchronize();
    k<<<148, nw * 32>>>(d, c, iters); cudaError_t e = cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    // each warp reads 16 KB per iteration; nw warps -> nw*16 KB per iteration per SM
    printf("%2d warps: %s  %.1f cycles/iter/warp  -> %.1f B/clk/SM ; one full 128x128 fp32 tile (64 KB) every %.0f cycles\n", nw, cudaGetErrorString(e),
           (double)h / iters, nw * 16384.0 * iters / h, 65536.0 / (nw * 16384.0 * iters / h));
  }
  return 0;
}
