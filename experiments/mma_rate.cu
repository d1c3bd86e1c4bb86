// Microbenchmark: tcgen05.mma issue-to-retire rate for the small tiles of the attention kernel.
//   one CTA per SM, one thread issues `n` MMAs back to back + one commit, waits on the mbarrier; cycles / n.
//   shapes: SS M128 N64 K16 (QK^T), TS M128 N48 K16 (PV), SS M128 N256 K16 (the GEMM's tile) for reference.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "vf_sm100.cuh"
using namespace vf::sm100;

__global__ void k(unsigned long long* cycles, int n_mma, int mode, int indep) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase;
  const int warp = threadIdx.x >> 5;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<512>(&tbase);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (threadIdx.x == 0) {
    const int n = mode == 2 ? 256 : (mode == 1 ? 48 : 64);
    const uint32_t idesc = mode == 1 ? make_idesc_bf16(128, 48, true) : make_idesc_bf16(128, n > 128 ? 128 : n, false);
    const uint64_t da = make_smem_desc_sw128(base, 16, 1024);
    const uint64_t db = mode == 1 ? make_smem_desc_sw128(base + 16384, 8192, 1024) : make_smem_desc_sw128(base + 16384, 16, 1024);
    unsigned long long t0 = clock64();
    for (int i = 0; i < n_mma; ++i) {
      const uint32_t acc_off = indep ? (uint32_t)(i % indep) * 0 : 0;   // placeholder
      if (mode == 1) mma_ts(tbase + 64 + (indep ? (i % indep) * 0 : 0), tbase + (i & 3) * 8, db, idesc, true);
      else if (indep) mma_ss(tbase + (uint32_t)(i % indep) * 64, da, db, make_idesc_bf16(128, 64, false), true);
      else mma_ss(tbase, da, db, idesc, true);
      (void)acc_off;
    }
    tc_commit(&bar);
    mbar_wait(&bar, 0);
    unsigned long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tbase); }
}

int main() {
  unsigned long long* cyc; cudaMalloc(&cyc, 148 * 4 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  const char* names[3] = {"SS M128 N64 K16 (QK^T)", "TS M128 N48 K16 (PV)", "SS M128 N128 K16"};
  for (int mode = 0; mode < 3; ++mode)
    for (int n : {1, 8, 64, 512}) {
      for (int rep = 0; rep < 2; ++rep) { k<<<148, 64, 48 * 1024>>>(cyc, n, mode, 0); cudaDeviceSynchronize(); }
      unsigned long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
      printf("%-26s n=%4d  %8llu clk total  %.1f clk per MMA   (%s)\n", names[mode], n, h[0], (double)h[0] / n,
             cudaGetErrorString(cudaGetLastError()));
    }
  for (int indep : {1, 2, 4, 8})
    for (int n : {64, 512}) {
      for (int rep = 0; rep < 2; ++rep) { k<<<148, 64, 48 * 1024>>>(cyc, n, 0, indep); cudaDeviceSynchronize(); }
      unsigned long long h[148]; cudaMemcpy(h, cyc, 148 * 8, cudaMemcpyDeviceToHost);
      printf("SS M128 N64 K16 round-robin over %d accumulators  n=%4d  %.1f clk per MMA\n", indep, n, (double)h[0] / n);
    }
  // two CTAs per SM issuing concurrently (independent accumulators in separate allocations)
  for (int n : {64, 512}) {
    for (int rep = 0; rep < 2; ++rep) { k<<<296, 64, 48 * 1024>>>(cyc, n, 0, 0); cudaDeviceSynchronize(); }
    unsigned long long h[296]; cudaMemcpy(h, cyc, 296 * 8, cudaMemcpyDeviceToHost);
    printf("SS M128 N64 K16, 2 CTAs/SM  n=%4d  %.1f clk per MMA per CTA (SM-wide %.1f)\n", n, (double)h[0] / n, (double)h[0] / n / 2);
  }
  return 0;
}
