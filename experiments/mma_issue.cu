// Microbenchmark: how fast can ONE warp issue small tcgen05.mma instructions?
//   A: the whole issue loop under `if (lane == 0)` (divergent: every descriptor goes R2UR right before the MMA)
//   B: warp-uniform loop, only the tcgen05.mma predicated on elect.sync (descriptors live in uniform registers)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vf_sm100.cuh"
using namespace vf::sm100;

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

template <int VARIANT>
__global__ void k(unsigned long long* cycles, int n_tiles) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<128>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t idesc_qk = make_idesc_bf16(128, 64, false);
    const uint32_t idesc_pv = make_idesc_bf16(128, 48, true);
    const uint32_t tm = tbase;
    unsigned long long t0 = clock64();
    if (VARIANT == 0) {
      if (lane == 0) {
        for (int j = 0; j < n_tiles; ++j) {
          const int st = j & 1;
          for (int s = 0; s < 3; ++s) {
            const uint64_t da = make_smem_desc_sw128(base + s * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(base + 16384 + st * 8192 + s * 32, 16, 1024);
            mma_ss(tm, da, db, idesc_qk, s > 0);
          }
          for (int s = 0; s < 4; ++s) {
            const uint64_t db = make_smem_desc_sw128(base + 32768 + st * 8192 + s * 2048, 8192, 1024);
            mma_ts(tm + 64, tm + 96 + s * 8, db, idesc_pv, (j > 0) || (s > 0));
          }
        }
        tc_commit(&bar);
      }
    } else {
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const uint64_t da = make_smem_desc_sw128(base + s * 32, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(base + 16384 + st * 8192 + s * 32, 16, 1024);
          if (elect_one()) mma_ss(tm, da, db, idesc_qk, s > 0);
        }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          const uint64_t db = make_smem_desc_sw128(base + 32768 + st * 8192 + s * 2048, 8192, 1024);
          if (elect_one()) mma_ts(tm + 64, tm + 96 + s * 8, db, idesc_pv, (j > 0) || (s > 0));
        }
      }
      if (elect_one()) tc_commit(&bar);
    }
    mbar_wait(&bar, 0);
    unsigned long long t1 = clock64();
    if (lane == 0) cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<128>(tbase); }
}

int main() {
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  for (int n : {16, 256}) {
    unsigned long long h[148];
    for (int rep = 0; rep < 2; ++rep) { k<0><<<148, 64, 56 * 1024>>>(cyc, n); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    printf("A (lane 0 branch)      tiles=%3d  %.1f clk per tile (7 MMAs) = %.1f clk per MMA   (%s)\n", n, (double)h[0] / n, (double)h[0] / n / 7, cudaGetErrorString(cudaGetLastError()));
    for (int rep = 0; rep < 2; ++rep) { k<1><<<148, 64, 56 * 1024>>>(cyc, n); cudaDeviceSynchronize(); }
    cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
    printf("B (uniform + elect)    tiles=%3d  %.1f clk per tile (7 MMAs) = %.1f clk per MMA   (%s)\n", n, (double)h[0] / n, (double)h[0] / n / 7, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
