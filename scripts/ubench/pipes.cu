// Throughput microbenchmarks for the instruction mix of the scoring epilogue (run on B200).
#include <cstdio>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#define ITER 4096
template <int MODE>
__global__ void k(float* out, float seed) {
  float a0 = seed + threadIdx.x, a1 = a0 * 0.5f, a2 = a0 * 0.25f, a3 = a0 * 0.125f;
  unsigned u0 = threadIdx.x, u1 = u0 + 1, u2 = u0 + 2, u3 = u0 + 3;
#pragma unroll 1
  for (int i = 0; i < ITER; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (MODE == 0) {  // MUFU.EX2 f32
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
        asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2)); asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
      } else if (MODE == 1) {  // F2FP bf16x2 pack
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u0) : "f"(a0), "f"(a1)); a0 = __uint_as_float(u0 << 16);
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u1) : "f"(a1), "f"(a2)); a1 = __uint_as_float(u1 << 16);
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u2) : "f"(a2), "f"(a3)); a2 = __uint_as_float(u2 << 16);
        asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(u3) : "f"(a3), "f"(a0)); a3 = __uint_as_float(u3 << 16);
      } else if (MODE == 2) {  // ex2 bf16x2
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u0)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u1));
        asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u2)); asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(u3));
      } else if (MODE == 3) {  // ex2 f16x2
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u0)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u1));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u2)); asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(u3));
      } else if (MODE == 4) {  // FFMA
        a0 = fmaf(a0, 1.0001f, 0.5f); a1 = fmaf(a1, 1.0001f, 0.5f); a2 = fmaf(a2, 1.0001f, 0.5f); a3 = fmaf(a3, 1.0001f, 0.5f);
      } else if (MODE == 5) {  // FFMA2
        float2 x = make_float2(a0, a1), y = make_float2(a2, a3), c = make_float2(1.0001f, 1.0001f), d = make_float2(.5f, .5f);
        x = __ffma2_rn(x, c, d); y = __ffma2_rn(y, c, d); x = __ffma2_rn(x, c, d); y = __ffma2_rn(y, c, d);
        a0 = x.x; a1 = x.y; a2 = y.x; a3 = y.y;
      } else if (MODE == 6) {  // F2FP f16x2 pack
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u0) : "f"(a0), "f"(a1)); a0 = __uint_as_float(u0 << 16);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u1) : "f"(a1), "f"(a2)); a1 = __uint_as_float(u1 << 16);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u2) : "f"(a2), "f"(a3)); a2 = __uint_as_float(u2 << 16);
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(u3) : "f"(a3), "f"(a0)); a3 = __uint_as_float(u3 << 16);
      } else if (MODE == 7) {  // packed bf16 max (HMNMX2)
        asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u0) : "r"(u1)); asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u1) : "r"(u2));
        asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u2) : "r"(u3)); asm volatile("max.bf16x2 %0, %0, %1;" : "+r"(u3) : "r"(u0));
      } else if (MODE == 8) {  // integer bf16 rounding emulation: (x + 0x7fff + ((x>>16)&1)) & 0xffff0000
        u0 = (u0 + 0x7fffu + ((u0 >> 16) & 1u)) & 0xffff0000u; u1 = (u1 + 0x7fffu + ((u1 >> 16) & 1u)) & 0xffff0000u;
        u2 = (u2 + 0x7fffu + ((u2 >> 16) & 1u)) & 0xffff0000u; u3 = (u3 + 0x7fffu + ((u3 >> 16) & 1u)) & 0xffff0000u;
      }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + __uint_as_float(u0 ^ u1 ^ u2 ^ u3);
}
template <int MODE>
void run(const char* name, int per_op_elems) {
  float* d; cudaMalloc(&d, 148 * 1024 * 4 * 8);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 2, 512>>>(d, 1.0f); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 2, 512>>>(d, 1.0f); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double ops = 148.0 * 2 * 512 * ITER * 8 * 4;  // thread-level instructions
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  double per_clk_sm = ops / (ms * 1e-3) / 148 / (clk * 1e3);
  printf("%-28s %8.3f ms  %7.1f thread-instr/clk/SM (at %d MHz nominal)  -> %.1f elements/clk/SM\n", name, ms, per_clk_sm, clk / 1000, per_clk_sm * per_op_elems);
  cudaFree(d);
}
int main() {
  run<0>("MUFU.EX2 f32", 1); run<1>("F2FP bf16x2 (+shift)", 2); run<6>("F2FP f16x2 (+shift)", 2); run<2>("ex2 bf16x2", 2); run<3>("ex2 f16x2", 2);
  run<4>("FFMA", 1); run<5>("FFMA2", 2); run<7>("HMNMX2 bf16x2", 2); run<8>("int bf16 round (4 ops)", 1);
  return 0;
}
