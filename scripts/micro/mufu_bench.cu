// Micro-benchmark: MUFU.EX2 issue rate per SM for f32, f16x2 and bf16x2 operands (is a packed ex2 one MUFU pass?).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mufu_bench mufu_bench.cu && ./mufu_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  uint32_t a[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (MODE == 0) ? __float_as_uint(-0.001f * (threadIdx.x + i + 1)) : 0x80108010u + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(a[i]));
      if (MODE == 1) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 2) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(a[i]));
      if (MODE == 3) {   // f32 pair -> pack -> packed ex2 (what a softmax pass would do)
        uint32_t p;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(p) : "f"(__uint_as_float(a[i])), "f"(__uint_as_float(a[(i + 1) & 7])));
        asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(p));
        a[i] ^= p & 1;
      }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int per_instr) {
  uint32_t* out;
  cudaMalloc(&out, 148 * 8 * 1024 * 4);
  const int iters = 4096;
  k<MODE><<<148 * 8, 1024>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  k<MODE><<<148 * 8, 1024>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double instr = 148.0 * 8 * 1024 * iters * 8;
  int clk;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-28s %.3f ms  %.1f lane-instr/clk/SM  (%.1f results/clk/SM at %d kHz)\n", name, ms,
         instr / (ms * 1e-3) / (clk * 1e3) / 148, instr * per_instr / (ms * 1e-3) / (clk * 1e3) / 148, clk);
  cudaFree(out);
}

int main() {
  run<0>("ex2.approx.ftz.f32", 1);
  run<1>("ex2.approx.f16x2", 2);
  run<2>("ex2.approx.ftz.bf16x2", 2);
  run<3>("cvt.f16x2 + ex2.f16x2", 2);
  return 0;
}
