// MUFU.EX2 throughput vs warps per SM scheduler, with the exact instruction pattern of the scoring epilogue.
#include <cstdio>
#include <cuda_runtime.h>
#define ITER 2048
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int MODE>
__global__ void k(float* out, const unsigned* in) {
  unsigned r[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) r[i] = in[threadIdx.x * 16 + i];
  float2 z = make_float2(0.f, 0.f), z1 = make_float2(0.f, 0.f);
  const float2 l2 = make_float2(1.44269504f, 1.44269504f), ml = make_float2(-3.f, -3.f);
#pragma unroll 1
  for (int it = 0; it < ITER; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float2 a = __ffma2_rn(make_float2(__uint_as_float(r[i] << 16), __uint_as_float(r[i] & 0xffff0000u)), l2, ml);
      if (MODE == 0) { z = __fadd2_rn(z, make_float2(ex2(a.x), ex2(a.y))); }
      if (MODE == 1) { if (i & 1) z = __fadd2_rn(z, make_float2(ex2(a.x), ex2(a.y))); else z1 = __fadd2_rn(z1, make_float2(ex2(a.x), ex2(a.y))); }
      if (MODE == 2) { z = __fadd2_rn(z, a); }  // no MUFU
      r[i] += it;
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = z.x + z.y + z1.x + z1.y;
}
template <int MODE>
void run(const char* name, int threads) {
  float* d; unsigned* in; cudaMalloc(&d, 148 * 1024 * 4); cudaMalloc(&in, 1024 * 16 * 4); cudaMemset(in, 0x3f, 1024 * 16 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(d, in); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, threads>>>(d, in); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double exps = 148.0 * threads * ITER * 32.0;
  printf("%-34s warps/SMSP=%4.1f  %7.3f ms  %6.2f exp/clk/SM (1.965 GHz)\n", name, threads / 128.0, ms, exps / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
  for (int t : {128, 256, 384, 512, 1024}) run<0>("unpack+FFMA2+2xMUFU+FADD2 (1 chain)", t);
  for (int t : {128, 256, 384, 512}) run<1>("same, 2 accumulation chains", t);
  for (int t : {256}) run<2>("no MUFU (FFMA2+FADD2 only)", t);
  return 0;
}
