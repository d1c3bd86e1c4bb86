// Microbenchmark: how does the sm_100a tensor pipe handle the SMALL tcgen05.mma instructions of the attention
// kernel when several CTAs per SM issue them, and does the ORDER (chains back to back, interleaved, SS vs TS,
// different N) matter?  One issuing warp per CTA (warp-uniform loop, elect.sync), R CTAs per SM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I vface_b200/csrc -o experiments/mma_mix.bin experiments/mma_mix.cu
#include <cstdio>
#include <cstdint>
#include <algorithm>
#include <cuda_runtime.h>
#include "vf_sm100.cuh"
using namespace vf::sm100;

struct Cfg { int main_cols; int bn; };

// PATTERN (MMAs per "tile"):
//  0: 7 x SS N64, one accumulator            1: 7 x TS N48, one accumulator
//  2: 3 SS N64 (S) then 4 TS N48 (O)         3: the same seven, interleaved S T S T S T T
//  4: 3 SS N64 (S) then 4 SS N48 MN-major (O) -- "P through shared memory"
//  5: 3 SS N64 then 4 TS N64 (same N)        6: CTA role: even resident index 7 x SS, odd 7 x TS
//  7: BN=128 tile: 3 SS N128 + 8 TS N48      8: BN=256 tile: 3 SS N256 + 16 TS N48
//  9: 3 SS N64 (S0) + 3 SS N64 (S1) + 4 TS (O0) + 4 TS (O1): two query tiles per CTA (14 MMAs)
template <int PATTERN>
__global__ void k(unsigned long long* cycles, unsigned* smids, int n_tiles) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tbase, tbase_p;
  constexpr bool kBig = PATTERN >= 7;               // one 512-column allocation, P at column 320, one CTA per SM
  constexpr int kMain = kBig ? 512 : 128;
  constexpr int kP = 32;
  constexpr int BN = PATTERN == 8 ? 256 : PATTERN == 7 ? 128 : 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) { tmem_alloc_only<kMain>(&tbase); if (!kBig) tmem_alloc_only<kP>(&tbase_p); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (warp == 0) {
    const uint32_t tm_s = tbase, tm_o = tbase + (PATTERN == 9 ? 128 : BN), tm_p = kBig ? tbase + 320 : tbase_p;
    const uint32_t id_qk = make_idesc_bf16(128, BN, false);
    const uint32_t id_pv = make_idesc_bf16(128, PATTERN == 5 ? 64 : 48, true);
    const uint32_t q_addr = base, k_addr = base + 16384, v_addr = k_addr + 2 * BN * 128;
    unsigned sm; asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    const int role = (blockIdx.x / 148) & 1;
    auto qk = [&](int st, int s, uint32_t acc) {
      const uint64_t da = make_smem_desc_sw128(q_addr + s * 32, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(k_addr + st * BN * 128 + s * 32, 16, 1024);
      if (elect_one()) mma_ss(acc, da, db, id_qk, s > 0);
    };
    auto pv = [&](int st, int s, uint32_t acc, uint32_t p, bool accum) {
      const uint64_t db = make_smem_desc_sw128(v_addr + st * BN * 128 + s * 2048, BN * 128, 1024);
      if (elect_one()) mma_ts(acc, p + s * 8, db, id_pv, accum);
    };
    auto pv_ss = [&](int st, int s, uint32_t acc, bool accum) {     // A = P tile in smem (K-major, 128 rows x 64 keys)
      const uint64_t da = make_smem_desc_sw128(q_addr + s * 32, 16, 1024);
      const uint64_t db = make_smem_desc_sw128(v_addr + st * BN * 128 + s * 2048, BN * 128, 1024);
      if (elect_one()) mma_ss(acc, da, db, id_pv, accum);
    };
    unsigned long long t0 = clock64();
    for (int j = 0; j < n_tiles; ++j) {
      const int st = j & 1;
      const bool acc = j > 0;
      if (PATTERN == 0 || (PATTERN == 6 && role == 0)) {
#pragma unroll
        for (int s = 0; s < 7; ++s) qk(st, s & 3, tm_s);
      } else if (PATTERN == 1 || (PATTERN == 6 && role == 1)) {
#pragma unroll
        for (int s = 0; s < 7; ++s) pv(st, s & 3, tm_o, tm_p, acc || s > 0);
      } else if (PATTERN == 2 || PATTERN == 5 || PATTERN == 7 || PATTERN == 8) {
#pragma unroll
        for (int s = 0; s < 3; ++s) qk(st, s, tm_s);
#pragma unroll
        for (int s = 0; s < BN / 16; ++s) pv(st, s, tm_o, tm_p, acc || s > 0);
      } else if (PATTERN == 3) {
#pragma unroll
        for (int s = 0; s < 4; ++s) {
          if (s < 3) qk(st, s, tm_s);
          pv(st, s, tm_o, tm_p, acc || s > 0);
        }
      } else if (PATTERN == 4) {
#pragma unroll
        for (int s = 0; s < 3; ++s) qk(st, s, tm_s);
#pragma unroll
        for (int s = 0; s < 4; ++s) pv_ss(st, s, tm_o, acc || s > 0);
      } else if (PATTERN == 9) {
#pragma unroll
        for (int s = 0; s < 3; ++s) qk(st, s, tm_s);
#pragma unroll
        for (int s = 0; s < 3; ++s) qk(st, s, tm_s + 64);
#pragma unroll
        for (int s = 0; s < 4; ++s) pv(st, s, tm_o, tm_p, acc || s > 0);
#pragma unroll
        for (int s = 0; s < 4; ++s) pv(st, s, tm_o + 64, tm_p + 32, acc || s > 0);
      }
    }
    if (elect_one()) tc_commit(&bar);
    mbar_wait(&bar, 0);
    unsigned long long t1 = clock64();
    if (lane == 0) { cycles[blockIdx.x] = t1 - t0; smids[blockIdx.x] = sm; }
  }
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc<kMain>(tbase); if (!kBig) tmem_dealloc<kP>(tbase_p); }
}

template <int PATTERN>
void run(const char* name, int mmas_per_tile, int elems_per_tile_k, int max_r, unsigned long long* cyc, unsigned* smid) {
  const int bn = PATTERN == 8 ? 256 : PATTERN == 7 ? 128 : 64;
  const size_t smem = 1024 + 16384 + 4 * (size_t)bn * 128;
  cudaFuncSetAttribute(k<PATTERN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  for (int r = 1; r <= max_r; ++r) {
    const int n = 256, grid = 148 * r;
    for (int rep = 0; rep < 2; ++rep) { k<PATTERN><<<grid, 32, smem>>>(cyc, smid, n); cudaDeviceSynchronize(); }
    static unsigned long long h[148 * 4]; static unsigned s[148 * 4];
    cudaMemcpy(h, cyc, grid * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(s, smid, grid * 4, cudaMemcpyDeviceToHost);
    int per_sm[256] = {0};
    for (int i = 0; i < grid; ++i) per_sm[s[i] & 255]++;
    int max_res = 0; for (int i = 0; i < 256; ++i) max_res = std::max(max_res, per_sm[i]);
    std::sort(h, h + grid);
    const double med = (double)h[grid / 2];
    printf("%-58s R=%d (max %d CTAs on an SM)  %7.1f clk/tile/CTA  %6.1f clk/MMA/CTA  SM-wide %6.1f clk/MMA  %6.3f clk per 1k S-elements  (%s)\n",
           name, r, max_res, med / n, med / n / mmas_per_tile, med / n / mmas_per_tile / r, med / n / r / elems_per_tile_k,
           cudaGetErrorString(cudaGetLastError()));
  }
}

int main() {
  unsigned long long* cyc; unsigned* smid;
  cudaMalloc(&cyc, 148 * 4 * 8); cudaMalloc(&smid, 148 * 4 * 4);
  run<0>("0: 7 x SS M128 N64 K16, one accumulator", 7, 8, 3, cyc, smid);
  run<1>("1: 7 x TS M128 N48 K16, one accumulator", 7, 8, 3, cyc, smid);
  run<2>("2: 3 SS N64 + 4 TS N48 (attention tile, chains back to back)", 7, 8, 3, cyc, smid);
  run<3>("3: same seven interleaved S T S T S T T", 7, 8, 3, cyc, smid);
  run<4>("4: 3 SS N64 + 4 SS N48 (P through smem)", 7, 8, 3, cyc, smid);
  run<5>("5: 3 SS N64 + 4 TS N64", 7, 8, 3, cyc, smid);
  run<6>("6: CTAs alternate roles: 7 SS | 7 TS", 7, 8, 3, cyc, smid);
  run<7>("7: BN=128 tile: 3 SS N128 + 8 TS N48", 11, 16, 1, cyc, smid);
  run<8>("8: BN=256 tile: 3 SS N256 + 16 TS N48", 19, 32, 1, cyc, smid);
  run<9>("9: two query tiles per CTA: 2x(3 SS N64) + 2x(4 TS N48)", 14, 16, 1, cyc, smid);
  return 0;
}
