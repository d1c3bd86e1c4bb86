// Feed-forward up-projection with the GEGLU gate fused into the GEMM epilogue (tcgen05 / TMEM / TMA).
//
//   out[r, 0:n] = (x[r] . Wv^T + bv) * gelu(x[r] . Wg^T + bg),   W = [Wv ; Wg]  ((2n, k) row-major, nn.Linear)
//
// Replaces GEGLU.forward (ldm/modules/attention.py:37-45: `x, gate = self.proj(x).chunk(2, dim=-1);
// return x * F.gelu(gate)`), i.e. a cuBLAS GEMM that writes the (rows, 2n) projection to HBM plus a
// memory-bound gate kernel that reads it back: at the 64x64 level (rows = 393 216, 2n = 2560) that is 2 GB
// written and 2 GB read per block for nothing.  Here the projection never leaves the SM: value and gate
// columns of an output tile are accumulated side by side in TMEM and gated on the way out.
//
// Kernel (persistent, one CTA per SM, warp-specialised):
//   warp 0      TMA producer: A tile 128 x 64 (rows x k) and two W boxes 128 x 64 (value rows n0.., gate rows
//               n + n0..) per k-block into a kStages ring, 128B swizzle, K-major both.
//   warp 1      TMEM allocator (all 512 columns) + tcgen05.mma issuer: per k-block four MMAs M128 N256 K16
//               (B = [Wv tile ; Wg tile] adjacent in shared memory -> accumulator columns [0,128) value,
//               [128,256) gate).  Two accumulator sets (2 x 256 columns) so that the epilogue of tile i
//               overlaps the MMAs of tile i+1.
//   warps 4..7  epilogue: tcgen05.ld 32 value + 32 gate columns per step (thread = row), + bias, exact-GELU
//               (A&S 7.1.26 erf, |err| < 5e-7), bf16x2 pack, 16-byte stores (each thread 64 contiguous bytes).
// Tiles are walked n-fastest, so the CTAs running at any instant share one or two A row-blocks through L2
// and W (a few MB) stays L2-resident.
#include "vf_common.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cstdlib>

namespace vf {

using namespace sm100;

constexpr int kGemmThreads = 256;
constexpr int kGemmBM = 128;          // rows per tile
constexpr int kGemmBN = 128;          // OUTPUT columns per tile (256 accumulator columns: value + gate)
constexpr int kGemmBK = 64;           // k-block: 64 bf16 = one 128-byte swizzled row
constexpr int kGemmStages = 3;
constexpr uint32_t kATileBytes = kGemmBM * kGemmBK * 2;          // 16 KB
constexpr uint32_t kBTileBytes = 2 * kGemmBN * kGemmBK * 2;      // 32 KB (value rows, then gate rows)
constexpr uint32_t kStageBytes = kATileBytes + kBTileBytes;

struct GemmGegluParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* bias;       // (2n) or null
  long long rows;
  int n, k;
  int m_blocks, n_blocks, k_blocks;
};

struct __align__(8) GemmBarriers {
  uint64_t full[kGemmStages], empty[kGemmStages];
  uint64_t acc_full[2], acc_empty[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// exact-GELU through the Abramowitz-Stegun 7.1.26 erf (same form as vf_norm.cu's bf16 GEGLU)
__device__ __forceinline__ float gelu_as(float g) {
  const float x = fabsf(g) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = p * t * exp2f(-1.4426950408889634f * x * x);
  return 0.5f * g * (1.0f + copysignf(1.0f - e, g));
}

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_geglu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const GemmGegluParams P) {
  extern __shared__ unsigned char gemm_smem[];
  __shared__ GemmBarriers bars;
  __shared__ __align__(16) float s_bias[2][2 * kGemmBN];   // per accumulator set: value bias [0,128), gate bias [128,256)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dyn_base = smem_u32(gemm_smem);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  unsigned char* tiles = gemm_smem + (tile_base - dyn_base);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kGemmStages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars.acc_full[a], 1);
      mbar_init(&bars.acc_empty[a], 4);            // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  const long long n_tiles = (long long)P.m_blocks * P.n_blocks;

  if (warp == 0) {
    // =========================== TMA producer ====================================================
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_w);
      uint32_t it = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mb = (int)(tile / P.n_blocks), nb = (int)(tile - (long long)mb * P.n_blocks);
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kGemmStages;
          const uint32_t use = it / kGemmStages;
          mbar_wait(&bars.empty[s], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.full[s], kStageBytes);
          unsigned char* st = tiles + (size_t)s * kStageBytes;
          tma_load_2d(st, &map_a, &bars.full[s], kb * kGemmBK, mb * kGemmBM);
          tma_load_2d(st + kATileBytes, &map_w, &bars.full[s], kb * kGemmBK, nb * kGemmBN);
          tma_load_2d(st + kATileBytes + kBTileBytes / 2, &map_w, &bars.full[s], kb * kGemmBK, P.n + nb * kGemmBN);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ======================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kGemmBM, 2 * kGemmBN, false);
      uint32_t it = 0, ti = 0;
      for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
        const uint32_t ab = ti & 1, ause = ti >> 1;
        mbar_wait(&bars.acc_empty[ab], (ause & 1) ^ 1);     // epilogue drained this accumulator set
        tc_fence_after();
        const uint32_t acc = tmem + ab * (2 * kGemmBN);
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kGemmStages;
          mbar_wait(&bars.full[s], (it / kGemmStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(tiles + (size_t)s * kStageBytes);
          const uint32_t b_addr = a_addr + kATileBytes;
#pragma unroll
          for (int ks = 0; ks < kGemmBK / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(a_addr + ks * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + ks * 32, 16, 1024);
            mma_ss(acc, da, db, idesc, (kb > 0) || (ks > 0));
          }
          tc_commit(&bars.empty[s]);                          // stage reusable once these MMAs retire
        }
        tc_commit(&bars.acc_full[ab]);
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue ========================================================
    const int quarter = warp & 3;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int et = threadIdx.x - 128;                        // 0..127
    uint32_t ti = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++ti) {
      const int mb = (int)(tile / P.n_blocks), nb = (int)(tile - (long long)mb * P.n_blocks);
      const uint32_t ab = ti & 1, ause = ti >> 1;
      // bias slice of this tile (double-buffered with the accumulator set: the previous user of s_bias[ab]
      // finished two tiles ago and every epilogue thread passed a named barrier since)
      s_bias[ab][et] = P.bias ? __bfloat162float(P.bias[nb * kGemmBN + et]) : 0.0f;
      s_bias[ab][kGemmBN + et] = P.bias ? __bfloat162float(P.bias[P.n + nb * kGemmBN + et]) : 0.0f;
      named_bar_sync(1, 128);
      mbar_wait(&bars.acc_full[ab], ause & 1);
      tc_fence_after();
      const uint32_t acc = tmem + ab * (2 * kGemmBN) + lane_off;
      const long long row = (long long)mb * kGemmBM + quarter * 32 + lane;
      __nv_bfloat16* orow = P.out + row * P.n + (long long)nb * kGemmBN;
#pragma unroll 1
      for (int c = 0; c < kGemmBN / 32; ++c) {
        uint32_t v[32], g[32];
        tmem_ld_x32(acc + c * 32, v);
        tmem_ld_x32(acc + kGemmBN + c * 32, g);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&s_bias[ab][c * 32 + i]);            // broadcast reads
          const float4 bg = *reinterpret_cast<const float4*>(&s_bias[ab][kGemmBN + c * 32 + i]);
          const float v0 = __uint_as_float(v[i]) + bv.x, v1 = __uint_as_float(v[i + 1]) + bv.y;
          const float v2 = __uint_as_float(v[i + 2]) + bv.z, v3 = __uint_as_float(v[i + 3]) + bv.w;
          const float g0 = __uint_as_float(g[i]) + bg.x, g1 = __uint_as_float(g[i + 1]) + bg.y;
          const float g2 = __uint_as_float(g[i + 2]) + bg.z, g3 = __uint_as_float(g[i + 3]) + bg.w;
          pk[i / 2] = pack_bf16(v0 * gelu_as(g0), v1 * gelu_as(g1));
          pk[i / 2 + 1] = pack_bf16(v2 * gelu_as(g2), v3 * gelu_as(g3));
        }
        if (row < P.rows) {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(orow + c * 32 + q * 8) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.acc_empty[ab]);
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

typedef CUresult (*EncodeTiledFn2)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn2 gemm_encode_fn() {
  static EncodeTiledFn2 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn2>(p);
  return fn;
}

// (rows, k) bf16 row-major, row stride ld elements -> 2-D map {k, rows}, box {64, 128}, 128B swizzle
static int make_map_2d(CUtensorMap* m, const void* base, long long rows, int k, long long ld, const char* who) {
  EncodeTiledFn2 enc = gemm_encode_fn();
  if (!enc) return fail("%s: cuTensorMapEncodeTiled entry point not found", who);
  cuuint64_t dims[2] = {(cuuint64_t)k, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)kGemmBK, (cuuint32_t)kGemmBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("%s: cuTensorMapEncodeTiled failed with CUresult %d", who, (int)r);
  return 0;
}

}  // namespace vf

extern "C" int vf_linear_geglu(const void* x, const void* w, const void* bias, void* out,
                               long long rows, int k, int n, long long ld_x, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !w || !out) return fail("vf_linear_geglu: null pointer");
  if (dtype != VF_BF16) return fail("vf_linear_geglu: only the bf16 path is fused (fp32 goes through the GEMM library + vf_geglu)");
  if (rows <= 0 || k <= 0 || n <= 0 || k % 8 || n % kGemmBN)
    return fail("vf_linear_geglu: bad shape rows=%lld k=%d n=%d (k %% 8 == 0, n %% %d == 0 required)", rows, k, n, kGemmBN);
  if (ld_x < k || ld_x % 8) return fail("vf_linear_geglu: bad row stride %lld", ld_x);
  const void* ptrs[3] = {x, w, out};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail("vf_linear_geglu: pointers must be 16-byte aligned");
  CUtensorMap ma, mw;
  if (int rc = make_map_2d(&ma, x, rows, k, ld_x, "vf_linear_geglu")) return rc;
  if (int rc = make_map_2d(&mw, w, 2LL * n, k, k, "vf_linear_geglu")) return rc;
  GemmGegluParams P;
  P.out = reinterpret_cast<__nv_bfloat16*>(out);
  P.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  P.rows = rows; P.n = n; P.k = k;
  P.m_blocks = (int)((rows + kGemmBM - 1) / kGemmBM);
  P.n_blocks = n / kGemmBN;
  P.k_blocks = (k + kGemmBK - 1) / kGemmBK;
  const size_t smem = 1024 + (size_t)kGemmStages * kStageBytes;
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  long long grid = (long long)P.m_blocks * P.n_blocks;
  if (grid > num_sms()) grid = num_sms();
  gemm_geglu_kernel<<<(int)grid, kGemmThreads, smem, (cudaStream_t)stream>>>(ma, mw, P);
  return check_cuda(cudaGetLastError(), "gemm_geglu_kernel launch");
}
