// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (TMEM <-> registers), 4..16 warps per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../vface_b200/csrc -o tmem_rate.bin tmem_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vf_sm100.cuh"
using namespace vf::sm100;

template <int MODE>   // 0: ld x32 + wait each; 1: 4 x ld x32 then wait; 2: st x32
__global__ void k(unsigned long long* cycles, uint32_t* sink, int iters) {
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t t = tbase + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = threadIdx.x + i;
  tmem_st_x32(t, r);
  tmem_wait_st();
  __syncthreads();
  unsigned long long t0 = clock64();
  uint32_t acc = 0;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      tmem_ld_x32(t + ((it & 3) * 32), r);
      tmem_wait_ld();
      acc ^= r[0] ^ r[13] ^ r[31];
    } else if (MODE == 1) {
      uint32_t a[32], b[32], c[32], d[32];
      tmem_ld_x32(t, a); tmem_ld_x32(t + 32, b);
      tmem_wait_ld();
      tmem_ld_x32(t + 64, c); tmem_ld_x32(t + 96, d);
      tmem_wait_ld();
      acc ^= a[0] ^ b[7] ^ c[19] ^ d[31] ^ a[31] ^ b[0] ^ c[1] ^ d[2];
    } else {
      r[0] = acc + it;
      tmem_st_x32(t + ((it & 3) * 32), r);
      tmem_wait_st();
      acc += 1;
    }
  }
  unsigned long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tbase);
}

template <int MODE>
void run(int warps, const char* name) {
  const int sms = 148, iters = 4000;
  unsigned long long* cyc; uint32_t* sink;
  cudaMalloc(&cyc, sms * 8); cudaMalloc(&sink, sms * warps * 32 * 4);
  k<MODE><<<sms, warps * 32>>>(cyc, sink, iters);
  cudaDeviceSynchronize();
  unsigned long long h[148];
  cudaMemcpy(h, cyc, sms * 8, cudaMemcpyDeviceToHost);
  double c = (double)h[0];
  const double per_iter_bytes = (MODE == 1 ? 4.0 : 1.0) * 32 * 32 * 4;      // per warp
  printf("%-26s warps/SM=%2d  %.1f clk per warp-iteration, %.1f B/clk per warp, %.1f B/clk per SM\n", name, warps,
         c / iters, per_iter_bytes * iters / c, per_iter_bytes * iters * warps / c);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  for (int w : {4, 8, 16}) run<0>(w, "ld.x32 + wait");
  for (int w : {4, 8, 16}) run<1>(w, "2 x (2 ld.x32 + wait)");
  for (int w : {4, 8, 16}) run<2>(w, "st.x32 + wait");
  return 0;
}
