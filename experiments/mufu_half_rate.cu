// Microbenchmark: are the 16-bit MUFU.EX2 variants (ex2.approx.f16x2 / ex2.approx.ftz.bf16x2) faster per element than
// the fp32 MUFU.EX2 on sm_100a?  (sm_103 doubles the SFU rate; sm_100 is not documented to.)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/mufu_half_rate.bin experiments/mufu_half_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(uint32_t* out, int iters) {
  uint32_t x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) x[i] = 0x3c003c00u + threadIdx.x + i;     // small positive halves
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { float f = __uint_as_float(x[i]); asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(f)); x[i] = __float_as_uint(f) & 0x3fffffffu; }
      if (MODE == 1) { asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i])); x[i] &= 0x3fff3fffu; }
      if (MODE == 2) { asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(x[i]) : "r"(x[i])); x[i] &= 0x3fff3fffu; }
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s ^= x[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(int warps, const char* name, int per_instr) {
  uint32_t* out; cudaMalloc(&out, 148 * warps * 32 * 4);
  const int iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, warps * 32>>>(out, 100); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148, warps * 32>>>(out, iters); cudaEventRecord(e1); cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  const double elems = (double)warps * 32 * iters * 8 * per_instr;
  printf("%-28s warps/SM=%2d  %.3f ms  %.2f exp2 results per clk per SM (at %.0f MHz max boost)\n", name, warps, ms,
         elems / (ms * 1e6) / (clk * 1e-6), clk * 1e-3);
  cudaFree(out);
}

int main() {
  for (int w : {8, 16, 32}) run<0>(w, "ex2.approx.ftz.f32", 1);
  for (int w : {8, 16, 32}) run<1>(w, "ex2.approx.f16x2", 2);
  for (int w : {8, 16, 32}) run<2>(w, "ex2.approx.ftz.bf16x2", 2);
  return 0;
}
