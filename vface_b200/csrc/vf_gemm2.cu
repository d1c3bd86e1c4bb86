// GEMM + GEGLU for k <= 320 on CTA PAIRS (tcgen05.mma.cta_group::2): the W-resident kernel of vf_gemm.cu with the
// value / gate weight tile split over the two SMs of a pair.
//
// Why.  The single-CTA W-resident kernel keeps the 160 KB value+gate weight tile of its n-block in shared memory, which
// leaves 48 KB for the A ring -- 48 KB in flight per SM against a loaded-L2 latency of ~1 us is ~17-25 B/clk/SM of the 32
// the tile's 2560 tensor cycles need (measured: 2 stages 0.614 ms, 3 stages 0.565 ms at the 64x64 level).  A CTA pair
// computes a 256-row tile with M = 256 MMAs: each CTA contributes its own 128 rows of A and HALF of the B operand -- CTA 0
// the value rows of the weight tile (accumulator columns [0,128)), CTA 1 the gate rows ([128,256)) -- so the resident
// weights cost 80 KB per SM and the A ring is eight stages (128 KB) deep.  Each CTA's TMEM ends up with its own 128 rows
// x 256 columns, and the epilogue (bias, exact GELU, gate, swizzled staging, TMA store) is the one of vf_gemm.cu.
//
// Protocol (rank 0 = leader):
//   full[s], w_full   live in the leader; BOTH CTAs' TMA loads complete_tx on them (cp.async.bulk.tensor ...
//                     .cta_group::2 with the barrier address' peer bit cleared), the leader's producer arms them
//   empty[s]          in each CTA, arrived by the leader's tcgen05.commit ... multicast::cluster (mask 0b11)
//   acc_full[a]       in each CTA, multicast commit after the last k-block of a tile
//   acc_empty[a]      in the leader, 2 x 8 arrivals: the epilogue warps of both CTAs (remote mbarrier.arrive from rank 1)
//   TMEM              512 columns allocated with tcgen05.alloc.cta_group::2 by warp 1 of both CTAs
#include "vf_common.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cstdlib>

namespace vf {

using namespace sm100;

constexpr int kG2Threads = 384;       // warp 0 TMA, 1 MMA (leader) + TMEM, 2-3 idle, 4-11 epilogue
constexpr int kG2BM = 128;            // rows per CTA (256 per pair)
constexpr int kG2BN = 128;            // OUTPUT columns per tile (256 accumulator columns: value + gate)
constexpr int kG2BK = 64;
constexpr int kG2Stages = 8;
constexpr uint32_t kG2ATile = kG2BM * kG2BK * 2;          // 16 KB
constexpr uint32_t kG2WTile = kG2BN * kG2BK * 2;          // 16 KB: this CTA's half (value OR gate rows) of a k-block
constexpr uint32_t kG2OutBytes = kG2BM * 32 * 2;          // 8 KB staging buffer (one 32-column step)
constexpr uint32_t kPeerMask = 0xFEFFFFFFu;               // clears the CTA-rank bit of a shared::cluster address -> rank 0

struct Gemm2Params {
  const __nv_bfloat16* bias;       // (2n) or null
  long long rows;
  int n, k;
  int m_blocks, n_blocks, k_blocks;       // m_blocks: 256-row blocks
};

struct __align__(8) Gemm2Barriers {
  uint64_t full[kG2Stages], empty[kG2Stages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// 2-D tile load into THIS CTA's shared memory, completion bytes to the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_2d_pair(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               :: "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :: "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate) : "memory");
}
// arrive on the mbarrier at the same shared-memory offset in BOTH CTAs once all MMAs issued so far have retired
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               :: "r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" :: "r"(smem_u32(bar) & kPeerMask) : "memory");
}
__device__ __forceinline__ void g2_named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// v * gelu(g) for a pair (A&S 7.1.26 erf, |err| < 5e-7; see vf_gemm.cu)
__device__ __forceinline__ float2 g2_splat(float c) { return make_float2(c, c); }
__device__ __forceinline__ float2 g2_geglu_pair(float2 v, float2 g) {
  const float2 a = make_float2(fabsf(g.x), fabsf(g.y));
  const float2 den = __ffma2_rn(a, g2_splat(0.3275911f * 0.70710678118654752f), g2_splat(1.0f));
  float2 t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(den.x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(den.y));
  float2 p = __ffma2_rn(g2_splat(1.061405429f), t, g2_splat(-1.453152027f));
  p = __ffma2_rn(p, t, g2_splat(1.421413741f));
  p = __ffma2_rn(p, t, g2_splat(-0.284496736f));
  p = __ffma2_rn(p, t, g2_splat(0.254829592f));
  const float2 xa = __fmul2_rn(__fmul2_rn(g, g), g2_splat(-0.5f * 1.4426950408889634f));
  float2 ex;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.x) : "f"(xa.x));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex.y) : "f"(xa.y));
  const float2 e = __fmul2_rn(__fmul2_rn(p, t), ex);
  const float2 m = __ffma2_rn(e, g2_splat(-1.0f), g2_splat(1.0f));
  const float2 sg = make_float2(copysignf(m.x, g.x), copysignf(m.y, g.y));
  const float2 hg = __fmul2_rn(g, g2_splat(0.5f));
  return __fmul2_rn(v, __ffma2_rn(hg, sg, hg));
}

// single-tanh form (see geglu_pair_tanh in vf_gemm.cu)
__device__ __forceinline__ float2 g2_geglu_pair_tanh(float2 v, float2 g) {
  const float2 gg = __fmul2_rn(g, g);
  const float2 y2 = make_float2(fminf(gg.x, 49.0f), fminf(gg.y, 49.0f));
  float2 p = __ffma2_rn(y2, g2_splat(-0.0003587323612f), g2_splat(0.0370503451f));
  p = __ffma2_rn(p, y2, g2_splat(0.7974584708f));
  const float2 u = __fmul2_rn(g, p);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hg = __fmul2_rn(g, g2_splat(0.5f));
  return __fmul2_rn(v, __ffma2_rn(hg, t, hg));
}

template <int kGelu>
__global__ void __launch_bounds__(kG2Threads, 1)
gemm_geglu_pair_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                       const __grid_constant__ CUtensorMap map_o, const Gemm2Params P) {
  //   [pad to 1024] | resident W half (k_blocks x 16 KB) | A ring (8 x 16 KB) | output staging (2 x 8 KB) | barriers | bias
  extern __shared__ unsigned char g2_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
  const uint32_t dyn_base = smem_u32(g2_smem);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  unsigned char* tiles = g2_smem + (tile_base - dyn_base);
  unsigned char* w_res = tiles;
  unsigned char* ring = tiles + (size_t)P.k_blocks * kG2WTile;
  unsigned char* out_stage = ring + (size_t)kG2Stages * kG2ATile;
  Gemm2Barriers& bars = *reinterpret_cast<Gemm2Barriers*>(out_stage + 2 * kG2OutBytes);
  float* s_bias = reinterpret_cast<float*>(out_stage + 2 * kG2OutBytes + 256);      // [256]: value bias, gate bias

  if (threadIdx.x == 0) {
    for (int s = 0; s < kG2Stages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    mbar_init(&bars.w_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars.acc_full[a], 1);
      mbar_init(&bars.acc_empty[a], 16);          // eight epilogue warps of each CTA (the leader's copy is the one used)
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"(smem_u32(&bars.tmem_base)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // both CTAs' barriers exist before anything signals across the pair
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  // the pair owns n-block nb and walks 256-row blocks group, group + groups, ...
  const int nb = pair % P.n_blocks;
  const int groups = pairs / P.n_blocks;
  const int mb0 = pair / P.n_blocks;

  if (warp == 0) {
    // =========================== TMA producer (both CTAs) =========================================
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_w);
      tma_prefetch_desc(&map_o);
      if (mb0 < P.m_blocks) {
        if (rank == 0) mbar_arrive_expect_tx(&bars.w_full, 2u * (uint32_t)P.k_blocks * kG2WTile);
        const int wrow = (rank == 0 ? 0 : P.n) + nb * kG2BN;            // value rows (rank 0) / gate rows (rank 1)
        for (int kb = 0; kb < P.k_blocks; ++kb)
          tma_load_2d_pair(w_res + (size_t)kb * kG2WTile, &map_w, &bars.w_full, kb * kG2BK, wrow);
      }
      uint32_t it = 0;
      for (int mb = mb0; mb < P.m_blocks; mb += groups) {
        const int row0 = mb * 2 * kG2BM + (int)rank * kG2BM;
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kG2Stages;
          mbar_wait(&bars.empty[s], ((it / kG2Stages) & 1) ^ 1);
          if (rank == 0) mbar_arrive_expect_tx(&bars.full[s], 2u * kG2ATile);
          tma_load_2d_pair(ring + (size_t)s * kG2ATile, &map_a, &bars.full[s], kb * kG2BK, row0);
          if (mb + groups < P.m_blocks) tma_prefetch_2d_pair(&map_a, kb * kG2BK, (mb + groups) * 2 * kG2BM + (int)rank * kG2BM);
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (leader CTA, one thread) ==============================
    if (rank == 0 && lane == 0) {
      const uint32_t idesc = make_idesc_bf16(2 * kG2BM, 2 * kG2BN, false);      // M = 256 over the pair, N = 256
      uint32_t it = 0, ti = 0;
      if (mb0 < P.m_blocks) {
        mbar_wait(&bars.w_full, 0);
        tc_fence_after();
      }
      for (int mb = mb0; mb < P.m_blocks; mb += groups, ++ti) {
        const uint32_t ab = ti & 1, ause = ti >> 1;
        mbar_wait(&bars.acc_empty[ab], (ause & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem + ab * (2 * kG2BN);
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kG2Stages;
          mbar_wait(&bars.full[s], (it / kG2Stages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)s * kG2ATile);
          const uint32_t b_addr = smem_u32(w_res + (size_t)kb * kG2WTile);
#pragma unroll
          for (int ks = 0; ks < kG2BK / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(a_addr + ks * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + ks * 32, 16, 1024);
            mma_ss_pair(acc, da, db, idesc, (kb > 0) || (ks > 0));
          }
          tc_commit_pair(&bars.empty[s]);
        }
        tc_commit_pair(&bars.acc_full[ab]);
      }
    }
  } else if (warp >= 4) {
    // =========================== epilogue (both CTAs, eight warps: two column halves) ==============
    const int quarter = warp & 3;
    const int ehalf = (warp - 4) >> 2;
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int et = threadIdx.x - 128;                      // 0..255
    if (et < kG2BN) {
      s_bias[et] = P.bias ? __bfloat162float(P.bias[nb * kG2BN + et]) : 0.0f;
      s_bias[kG2BN + et] = P.bias ? __bfloat162float(P.bias[P.n + nb * kG2BN + et]) : 0.0f;
    }
    g2_named_bar_sync(1, 256);
    const int trow = quarter * 32 + lane;
    uint32_t ti = 0;
    for (int mb = mb0; mb < P.m_blocks; mb += groups, ++ti) {
      const uint32_t ab = ti & 1, ause = ti >> 1;
      mbar_wait(&bars.acc_full[ab], ause & 1);
      tc_fence_after();
      const uint32_t acc = tmem + ab * (2 * kG2BN) + lane_off;
      const int row0 = mb * 2 * kG2BM + (int)rank * kG2BM;
      unsigned char* obuf = out_stage + (size_t)ehalf * kG2OutBytes;
      const uint32_t sw = (uint32_t)(trow >> 1) & 3u;
      const bool issuer = (threadIdx.x & 127) == 0;
#pragma unroll 1
      for (int c = ehalf * 2; c < ehalf * 2 + 2; ++c) {
        uint32_t v[32], g[32];
        tmem_ld_x32(acc + c * 32, v);
        tmem_ld_x32(acc + kG2BN + c * 32, g);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&s_bias[c * 32 + i]);
          const float4 bg = *reinterpret_cast<const float4*>(&s_bias[kG2BN + c * 32 + i]);
          const float2 va = __fadd2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(bv.x, bv.y));
          const float2 vb = __fadd2_rn(make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), make_float2(bv.z, bv.w));
          const float2 ga = __fadd2_rn(make_float2(__uint_as_float(g[i]), __uint_as_float(g[i + 1])), make_float2(bg.x, bg.y));
          const float2 gb = __fadd2_rn(make_float2(__uint_as_float(g[i + 2]), __uint_as_float(g[i + 3])), make_float2(bg.z, bg.w));
          const float2 oa = kGelu ? g2_geglu_pair_tanh(va, ga) : g2_geglu_pair(va, ga);
          const float2 ob = kGelu ? g2_geglu_pair_tanh(vb, gb) : g2_geglu_pair(vb, gb);
          pk[i / 2] = pack_bf16(oa.x, oa.y);
          pk[i / 2 + 1] = pack_bf16(ob.x, ob.y);
        }
        // single staging buffer per group: the previous step's store must have drained it
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        g2_named_bar_sync(2 + ehalf, 128);
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(obuf + trow * 64 + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async();
        g2_named_bar_sync(2 + ehalf, 128);
        if (issuer) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       :: "l"(reinterpret_cast<uint64_t>(&map_o)), "r"(nb * kG2BN + c * 32), "r"(row0), "r"(smem_u32(obuf))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars.acc_empty[ab]);
    }
    if ((threadIdx.x & 127) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                              // neither CTA frees TMEM / exits while the other may still signal it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" :: "r"(tmem) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int g2_make_map(CUtensorMap* m, const void* base, long long rows, int cols, long long ld, int box_cols, int box_rows,
                       CUtensorMapSwizzle sw) {
  static EncodeTiledFn3 enc = nullptr;
  if (!enc) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
      return fail("vf_linear_geglu: cuTensorMapEncodeTiled entry point not found");
    enc = reinterpret_cast<EncodeTiledFn3>(p);
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("vf_linear_geglu: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// -1: shape not eligible (caller falls back to the single-CTA kernels); 0 / 1: launched / error
int launch_gemm_geglu_pair(const void* x, const void* w, const void* bias, void* out, long long rows, int k, int n,
                           long long ld_x, cudaStream_t st) {
  Gemm2Params P;
  P.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  P.rows = rows; P.n = n; P.k = k;
  P.m_blocks = (int)((rows + 2 * kG2BM - 1) / (2 * kG2BM));
  P.n_blocks = n / kG2BN;
  P.k_blocks = (k + kG2BK - 1) / kG2BK;
  const size_t smem = 1008 + (size_t)P.k_blocks * kG2WTile + (size_t)kG2Stages * kG2ATile + 2 * kG2OutBytes + 256 + 2 * kG2BN * sizeof(float);
  const int pairs_avail = num_sms() / 2;
  if (smem > 232448 || P.n_blocks > pairs_avail || P.m_blocks < 4 * (pairs_avail / P.n_blocks)) return -1;
  CUtensorMap ma, mw, mo;
  if (int rc = g2_make_map(&ma, x, rows, k, ld_x, kG2BK, kG2BM, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = g2_make_map(&mw, w, 2LL * n, k, k, kG2BK, kG2BN, CU_TENSOR_MAP_SWIZZLE_128B)) return rc;
  if (int rc = g2_make_map(&mo, out, rows, n, n, 32, kG2BM, CU_TENSOR_MAP_SWIZZLE_64B)) return rc;
  static int gelu_knob = -1;
  if (gelu_knob < 0) { const char* e = getenv("VF_GEMM_GELU"); gelu_knob = e ? atoi(e) : 1; }
  static size_t attr = 0;
  if (smem > attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_pair_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_pair_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const int groups = pairs_avail / P.n_blocks;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(2 * groups * P.n_blocks), 1, 1);
  cfg.blockDim = dim3(kG2Threads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = 2;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return check_cuda(gelu_knob ? cudaLaunchKernelEx(&cfg, gemm_geglu_pair_kernel<1>, ma, mw, mo, P)
                              : cudaLaunchKernelEx(&cfg, gemm_geglu_pair_kernel<0>, ma, mw, mo, P), "gemm_geglu_pair_kernel launch");
}

}  // namespace vf
