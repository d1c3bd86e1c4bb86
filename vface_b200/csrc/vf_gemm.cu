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
//
// kWResident (k <= 320, the 64x64 level): at k = 320 a tile moves 240 KB through L2 for 2560 tensor cycles of
// work -- 94 B/clk/SM against the ~42 B/clk/SM that the L2 can feed 148 SMs -- and the streaming kernel is
// L2-bound (measured 8.5 TB/s of L2 reads).  The value+gate weight tile of ONE n-block for ALL of k is only
// 160 KB, so each CTA keeps it in shared memory for its whole life, walks the m-blocks of its n-block and
// streams only A (80 KB per tile).
#include "vf_common.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cstdlib>

namespace vf {

using namespace sm100;

constexpr int kGemmThreads = 384;     // warp 0 TMA, 1 MMA, 2-3 idle, 4-7 epilogue, 8-11 second epilogue group (streaming kernel only)
constexpr int kGemmBM = 128;          // rows per tile
constexpr int kGemmBN = 128;          // OUTPUT columns per tile (256 accumulator columns: value + gate)
constexpr int kGemmBK = 64;           // k-block: 64 bf16 = one 128-byte swizzled row
constexpr int kGemmStages = 4;        // ring depth of the streaming kernel (A + W per stage)
#ifndef VF_GEMM_STAGES_RES
#define VF_GEMM_STAGES_RES 3
#endif
constexpr int kGemmStagesRes = VF_GEMM_STAGES_RES;     // ring depth of the W-resident kernel (A only): the fourth stage's 16 KB hold the output staging
constexpr uint32_t kATileBytes = kGemmBM * kGemmBK * 2;          // 16 KB
constexpr uint32_t kBTileBytes = 2 * kGemmBN * kGemmBK * 2;      // 32 KB (value rows, then gate rows)
constexpr uint32_t kStageBytes = kATileBytes + kBTileBytes;
// Output staging: one column step (128 rows x 32 bf16 = 64-byte rows, 64B-swizzled) per buffer, written by the
// epilogue threads and drained by a TMA store.  A thread owns a ROW of the accumulator (TMEM lane), so storing
// straight from registers made every 16-byte store instruction touch 32 different rows: 2048 l1tex wavefronts per
// tile, 66 % of the LSU data pipe (profiles/r1_geglu_gemm_ncu.txt) -- the same order as the tile's 2560 tensor cycles.
// Through the staging buffer a warp's store is 4 conflict-free shared-memory wavefronts and the global write is the
// TMA engine's.
constexpr uint32_t kStageOutBytes = kGemmBM * 32 * 2;            // 8 KB

struct GemmGegluParams {
  __nv_bfloat16* out;
  const __nv_bfloat16* bias;       // (2n) or null
  long long rows;
  int n, k;
  int m_blocks, n_blocks, k_blocks;
  int prefetch;                    // W-resident: L2 prefetch distance of one tile (VF_GEMM_PREFETCH, default on)
};

struct __align__(8) GemmBarriers {
  uint64_t full[kGemmStages], empty[kGemmStages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t w_full;
  uint32_t tmem_base;
};

// tile walk.  streaming: tile = blockIdx.x + i*gridDim.x, n-fastest.  W-resident: the CTA owns n-block
// blockIdx.x % n_blocks and walks m-blocks group, group + groups, ... (groups = gridDim.x / n_blocks).
template <bool kWResident>
struct TileWalk {
  long long cur, end, step;
  int nb_fixed, n_blocks;
  __device__ TileWalk(const GemmGegluParams& P) {
    n_blocks = P.n_blocks;
    if (kWResident) {
      nb_fixed = blockIdx.x % P.n_blocks;
      cur = blockIdx.x / P.n_blocks;
      step = gridDim.x / P.n_blocks;
      end = P.m_blocks;
    } else {
      nb_fixed = 0;
      cur = blockIdx.x;
      step = gridDim.x;
      end = (long long)P.m_blocks * P.n_blocks;
    }
  }
  __device__ bool valid() const { return cur < end; }
  __device__ void next() { cur += step; }
  __device__ int mb() const { return kWResident ? (int)cur : (int)(cur / n_blocks); }
  __device__ int nb() const { return kWResident ? nb_fixed : (int)(cur % n_blocks); }
};

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// L2 prefetch of a tile (no shared-memory destination, no completion tracking)
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               :: "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// v * gelu(g) for a PAIR of elements, exact (erf) GELU through Abramowitz-Stegun 7.1.26 (|err| < 5e-7), written
// for the epilogue's issue budget: every fp32 operation is a packed FFMA2/FMUL2/FADD2 over the pair, the two
// transcendental steps are one MUFU each (rcp.approx, ex2.approx), |g| and copysign are single LOP3s:
// ~12 issue slots per element against 27 for the scalar libdevice form (which made the epilogue, not the
// tensor pipe, the bound at k = 320).
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float2 splat2(float c) { return make_float2(c, c); }

__device__ __forceinline__ float2 geglu_pair(float2 v, float2 g) {
  const float2 a = make_float2(fabsf(g.x), fabsf(g.y));
  const float2 den = __ffma2_rn(a, splat2(0.3275911f * 0.70710678118654752f), splat2(1.0f));
  const float2 t = make_float2(rcp_approx(den.x), rcp_approx(den.y));
  float2 p = __ffma2_rn(splat2(1.061405429f), t, splat2(-1.453152027f));
  p = __ffma2_rn(p, t, splat2(1.421413741f));
  p = __ffma2_rn(p, t, splat2(-0.284496736f));
  p = __ffma2_rn(p, t, splat2(0.254829592f));
  const float2 xa = __fmul2_rn(__fmul2_rn(g, g), splat2(-0.5f * 1.4426950408889634f));     // -(g/sqrt2)^2 * log2(e)
  const float2 ex = make_float2(ex2_approx_ftz(xa.x), ex2_approx_ftz(xa.y));
  const float2 e = __fmul2_rn(__fmul2_rn(p, t), ex);                                        // 1 - erf(|g|/sqrt2)
  const float2 m = __ffma2_rn(e, splat2(-1.0f), splat2(1.0f));                              // erf(|g|/sqrt2)
  const float2 sg = make_float2(copysignf(m.x, g.x), copysignf(m.y, g.y));
  const float2 hg = __fmul2_rn(g, splat2(0.5f));
  return __fmul2_rn(v, __ffma2_rn(hg, sg, hg));
}

// The same product with ONE transcendental per element: Phi(g) = 1/2 + 1/2 tanh(u), u = g (c1 + c3 g^2 + c5 g^4) fitted to
// atanh(erf(g / sqrt 2)) (max |error| of g Phi(g) 3e-5 in exact arithmetic, g^2 clamped at 49 so that the quintic cannot
// turn over: beyond |g| = 7 the tanh is saturated either way), evaluated with tanh.approx.f32 (|rel err| < 2^-10.9, the
// instruction the GroupNorm's SiLU already uses): ~2.5e-4 |g| absolute, an eighth of the bf16 rounding of the output.
// The A&S form costs two MUFU (rcp, ex2) and ~12 issue slots per element, and at k = 320 the EPILOGUE bounds the kernel
// (a CTA pair with an eight-stage A ring runs at the same 0.55 ms as the single CTA with three: vf_gemm2.cu); this one
// costs one MUFU and ~6 slots.  VF_GEMM_GELU=0 selects the A&S form.
__device__ __forceinline__ float2 geglu_pair_tanh(float2 v, float2 g) {
  const float2 g2 = __fmul2_rn(g, g);
  const float2 y2 = make_float2(fminf(g2.x, 49.0f), fminf(g2.y, 49.0f));
  float2 p = __ffma2_rn(y2, splat2(-0.0003587323612f), splat2(0.0370503451f));
  p = __ffma2_rn(p, y2, splat2(0.7974584708f));
  const float2 u = __fmul2_rn(g, p);
  float2 t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.x) : "f"(u.x));
  asm("tanh.approx.f32 %0, %1;" : "=f"(t.y) : "f"(u.y));
  const float2 hg = __fmul2_rn(g, splat2(0.5f));
  return __fmul2_rn(v, __ffma2_rn(hg, t, hg));               // v * (g/2 + g/2 tanh u)
}

template <bool kWResident, int kEpiGroups, int kGelu = 1>      // kEpiGroups: 1 = four epilogue warps (4-7), 2 = eight (4-11); kGelu: 0 A&S erf, 1 single tanh
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_geglu_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                  const __grid_constant__ CUtensorMap map_o, const GemmGegluParams P) {
  // all shared memory is dynamic (the W-resident layout uses the 227 KB to within 1 KB):
  //   [pad to 1024] | resident W (k_blocks x 32 KB, W-resident only) | ring | output staging | barriers | bias
  extern __shared__ unsigned char gemm_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dyn_base = smem_u32(gemm_smem);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  unsigned char* tiles = gemm_smem + (tile_base - dyn_base);
  constexpr int kStages = kWResident ? kGemmStagesRes : kGemmStages;
  constexpr int kOutBufs = (kWResident && kEpiGroups == 1) ? 2 : 1;     // staging buffers per epilogue group: 16 KB in all
  constexpr int kEpiGroupsC = kEpiGroups;
  const size_t ring_bytes = (kWResident ? (size_t)P.k_blocks * kBTileBytes + (size_t)kStages * kATileBytes
                                        : (size_t)kStages * kStageBytes);
  unsigned char* out_stage = tiles + ring_bytes;          // 1024-aligned: every part before it is a multiple of 1 KB
  const size_t tile_bytes = ring_bytes + (size_t)kEpiGroupsC * kOutBufs * kStageOutBytes;
  GemmBarriers& bars = *reinterpret_cast<GemmBarriers*>(tiles + tile_bytes);
  // per accumulator set: value bias [0,128), gate bias [128,256)
  float (*s_bias)[2 * kGemmBN] = reinterpret_cast<float (*)[2 * kGemmBN]>(tiles + tile_bytes + 128);

  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], 1);
    }
    mbar_init(&bars.w_full, 1);
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars.acc_full[a], 1);
      mbar_init(&bars.acc_empty[a], 4 * kEpiGroups);                // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  // shared memory: [resident W: k_blocks x 32 KB] | A(+W) ring
  unsigned char* w_res = tiles;
  unsigned char* ring = kWResident ? tiles + (size_t)P.k_blocks * kBTileBytes : tiles;
  constexpr uint32_t kRingStage = kWResident ? kATileBytes : kStageBytes;

  if (warp == 0) {
    // =========================== TMA producer ====================================================
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_w);
      tma_prefetch_desc(&map_o);
      uint32_t it = 0;
      TileWalk<kWResident> tw(P);
      if (kWResident && tw.valid()) {
        mbar_arrive_expect_tx(&bars.w_full, (uint32_t)P.k_blocks * kBTileBytes);
        for (int kb = 0; kb < P.k_blocks; ++kb) {
          unsigned char* wd = w_res + (size_t)kb * kBTileBytes;
          tma_load_2d(wd, &map_w, &bars.w_full, kb * kGemmBK, tw.nb() * kGemmBN);
          tma_load_2d(wd + kBTileBytes / 2, &map_w, &bars.w_full, kb * kGemmBK, P.n + tw.nb() * kGemmBN);
        }
      }
      for (; tw.valid(); tw.next()) {
        const int mb = tw.mb(), nb = tw.nb();
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          const uint32_t use = it / kStages;
          mbar_wait(&bars.empty[s], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.full[s], kRingStage);
          unsigned char* st = ring + (size_t)s * kRingStage;
          tma_load_2d(st, &map_a, &bars.full[s], kb * kGemmBK, mb * kGemmBM);
          // W-resident: the ring holds 48 KB per SM, i.e. 48 KB per memory latency -- with A coming from HBM (~1.2 us
          // under load) that is ~17 B/clk/SM, half of what the tile's 2560 tensor cycles need.  Pull the same k-block
          // of the NEXT tile into L2 now, so that its load is an L2 hit when its turn comes.
          if (kWResident && (P.prefetch & 1) && tw.cur + tw.step < tw.end)
            tma_prefetch_2d(&map_a, kb * kGemmBK, (int)(tw.cur + tw.step) * kGemmBM);
          if (!kWResident && (P.prefetch & 2) && tw.cur + tw.step < tw.end)            // streaming: next tile's A row block
            tma_prefetch_2d(&map_a, kb * kGemmBK, (int)((tw.cur + tw.step) / tw.n_blocks) * kGemmBM);
          if (!kWResident) {
            tma_load_2d(st + kATileBytes, &map_w, &bars.full[s], kb * kGemmBK, nb * kGemmBN);
            tma_load_2d(st + kATileBytes + kBTileBytes / 2, &map_w, &bars.full[s], kb * kGemmBK, P.n + nb * kGemmBN);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ======================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kGemmBM, 2 * kGemmBN, false);
      uint32_t it = 0, ti = 0;
      TileWalk<kWResident> tw(P);
      if (kWResident && tw.valid()) {
        mbar_wait(&bars.w_full, 0);
        tc_fence_after();
      }
      for (; tw.valid(); tw.next(), ++ti) {
        const uint32_t ab = ti & 1, ause = ti >> 1;
        mbar_wait(&bars.acc_empty[ab], (ause & 1) ^ 1);     // epilogue drained this accumulator set
        tc_fence_after();
        const uint32_t acc = tmem + ab * (2 * kGemmBN);
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(&bars.full[s], (it / kStages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)s * kRingStage);
          const uint32_t b_addr = kWResident ? smem_u32(w_res + (size_t)kb * kBTileBytes) : a_addr + kATileBytes;
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
  } else if (warp >= 4 && warp < 4 + 4 * kEpiGroups) {
    // =========================== epilogue ========================================================
    // Eight epilogue warps: warps 4-7 and 8-11 both cover the four TMEM lane quarters (a warp may touch quarter
    // warp % 4) and split the column steps of every tile.  With four, ONE warp per scheduler had to push the ~12
    // issue slots per output element at single-warp IPC.  Measured: k = 640 0.567 -> 0.539 ms, k = 1280 0.480 -> 0.465;
    // the W-resident kernel (k = 320) got SLOWER with eight (0.659 -> 0.739 ms) and keeps four.
    const int quarter = warp & 3;
    const int ehalf = (warp - 4) >> 2;                       // group 0: the first half of the column steps, 1: the rest
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int et = threadIdx.x - 128;                        // 0..255; the first 128 stage the bias
    uint32_t ti = 0, ostep = 0;
    for (TileWalk<kWResident> tw(P); tw.valid(); tw.next(), ++ti) {
      const int mb = tw.mb(), nb = tw.nb();
      const uint32_t ab = ti & 1, ause = ti >> 1;
      // bias slice of this tile (double-buffered with the accumulator set: the previous user of s_bias[ab]
      // finished two tiles ago and every epilogue thread passed a named barrier since)
      // (W-resident: the n-block never changes, one set filled on the first tile)
      const uint32_t bs = kWResident ? 0u : ab;
      if (!kWResident || ti == 0) {
        if (et < kGemmBN) {
          s_bias[bs][et] = P.bias ? __bfloat162float(P.bias[nb * kGemmBN + et]) : 0.0f;
          s_bias[bs][kGemmBN + et] = P.bias ? __bfloat162float(P.bias[P.n + nb * kGemmBN + et]) : 0.0f;
        }
        named_bar_sync(1, 128 * kEpiGroups);
      }
      mbar_wait(&bars.acc_full[ab], ause & 1);
      tc_fence_after();
      const uint32_t acc = tmem + ab * (2 * kGemmBN) + lane_off;
      const int trow = quarter * 32 + lane;                  // row of the tile this thread owns
#pragma unroll 1
      for (int c = ehalf * (kGemmBN / 32 / kEpiGroups); c < (ehalf + 1) * (kGemmBN / 32 / kEpiGroups); ++c, ++ostep) {
        uint32_t v[32], g[32];
        tmem_ld_x32(acc + c * 32, v);
        tmem_ld_x32(acc + kGemmBN + c * 32, g);
        tmem_wait_ld();
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          const float4 bv = *reinterpret_cast<const float4*>(&s_bias[bs][c * 32 + i]);            // broadcast reads
          const float4 bg = *reinterpret_cast<const float4*>(&s_bias[bs][kGemmBN + c * 32 + i]);
          const float2 va = __fadd2_rn(make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1])), make_float2(bv.x, bv.y));
          const float2 vb = __fadd2_rn(make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3])), make_float2(bv.z, bv.w));
          const float2 ga = __fadd2_rn(make_float2(__uint_as_float(g[i]), __uint_as_float(g[i + 1])), make_float2(bg.x, bg.y));
          const float2 gb = __fadd2_rn(make_float2(__uint_as_float(g[i + 2]), __uint_as_float(g[i + 3])), make_float2(bg.z, bg.w));
          const float2 oa = kGelu ? geglu_pair_tanh(va, ga) : geglu_pair(va, ga);
          const float2 ob = kGelu ? geglu_pair_tanh(vb, gb) : geglu_pair(vb, gb);
          pk[i / 2] = pack_bf16(oa.x, oa.y);
          pk[i / 2 + 1] = pack_bf16(ob.x, ob.y);
        }
        // 64-byte row of this step into the staging buffer, TMA's 64B swizzle: 16-byte chunk q of row r sits at
        // chunk q ^ ((r >> 1) & 3) -- eight consecutive rows cover all 32 banks, 4 wavefronts per warp instruction.
        unsigned char* obuf = out_stage + (size_t)(ehalf * kOutBufs + (kOutBufs == 2 ? (ostep & 1) : 0)) * kStageOutBytes;
        const uint32_t sw = (uint32_t)(trow >> 1) & 3u;
        const bool issuer = (threadIdx.x & 127) == 0;
        if (kOutBufs == 1) {                                   // single buffer: the previous step's store must have drained it
          if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          named_bar_sync(2 + ehalf, 128);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(obuf + trow * 64 + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        fence_proxy_async();
        // the issuing thread first makes sure the OTHER buffer (written next) has been drained by its store, then the
        // group meets, then this buffer goes out: one barrier per step
        if (kOutBufs == 2 && issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        named_bar_sync(2 + ehalf, 128);
        if (issuer) {
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       :: "l"(reinterpret_cast<uint64_t>(&map_o)), "r"(nb * kGemmBN + c * 32), "r"(mb * kGemmBM), "r"(smem_u32(obuf))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.acc_empty[ab]);
    }
    if ((threadIdx.x & 127) == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");     // stores complete before exit
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

// output (rows, n) bf16 row-major -> 2-D map {n, rows}, box {32, 128}, 64B swizzle (the staging layout of the epilogue)
static int make_map_out(CUtensorMap* m, void* base, long long rows, int n, const char* who) {
  EncodeTiledFn2 enc = gemm_encode_fn();
  if (!enc) return fail("%s: cuTensorMapEncodeTiled entry point not found", who);
  cuuint64_t dims[2] = {(cuuint64_t)n, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)n * 2};
  cuuint32_t box[2] = {32, (cuuint32_t)kGemmBM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("%s: cuTensorMapEncodeTiled (output) failed with CUresult %d", who, (int)r);
  return 0;
}

// CTA-pair kernel of vf_gemm2.cu (k <= 320): -1 = shape not eligible
int launch_gemm_geglu_pair(const void* x, const void* w, const void* bias, void* out, long long rows, int k, int n,
                           long long ld_x, cudaStream_t st);

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
  // VF_GEMM_PAIR=1: CTA pairs (tcgen05.mma.cta_group::2, vf_gemm2.cu) where the weight tile would otherwise be resident
  static int pair_knob = -1;
  if (pair_knob < 0) { const char* e = getenv("VF_GEMM_PAIR"); pair_knob = e ? atoi(e) : 0; }
  if (pair_knob && k <= 320 && k % 64 == 0) {
    const int rc = launch_gemm_geglu_pair(x, w, bias, out, rows, k, n, ld_x, (cudaStream_t)stream);
    if (rc >= 0) return rc;
  }
  CUtensorMap ma, mw, mo;
  if (int rc = make_map_2d(&ma, x, rows, k, ld_x, "vf_linear_geglu")) return rc;
  if (int rc = make_map_2d(&mw, w, 2LL * n, k, k, "vf_linear_geglu")) return rc;
  if (int rc = make_map_out(&mo, out, rows, n, "vf_linear_geglu")) return rc;
  GemmGegluParams P;
  P.out = reinterpret_cast<__nv_bfloat16*>(out);
  P.bias = reinterpret_cast<const __nv_bfloat16*>(bias);
  P.rows = rows; P.n = n; P.k = k;
  P.m_blocks = (int)((rows + kGemmBM - 1) / kGemmBM);
  P.n_blocks = n / kGemmBN;
  P.k_blocks = (k + kGemmBK - 1) / kGemmBK;
  static int w_res_knob = -1;      // VF_GEMM_WRES=0 forces the streaming kernel
  if (w_res_knob < 0) { const char* e = getenv("VF_GEMM_WRES"); w_res_knob = e ? atoi(e) : 1; }
  constexpr size_t kTail = 128 + 2 * 2 * kGemmBN * sizeof(float);      // barriers + bias
  static_assert(sizeof(GemmBarriers) <= 128, "barrier block");
  static int pf_knob = -1;
  if (pf_knob < 0) { const char* e = getenv("VF_GEMM_PREFETCH"); pf_knob = e ? atoi(e) : 1; }
  P.prefetch = pf_knob;
  static int gelu_knob = -1;       // VF_GEMM_GELU: 1 (default) single-tanh GELU in the epilogue, 0 the A&S erf form
  if (gelu_knob < 0) { const char* e = getenv("VF_GEMM_GELU"); gelu_knob = e ? atoi(e) : 1; }
  static int epi_res = -1;         // VF_GEMM_EPI_RES: epilogue warp groups of the W-resident kernel (1 or 2)
  if (epi_res < 0) { const char* e = getenv("VF_GEMM_EPI_RES"); epi_res = e ? atoi(e) : 2; if (epi_res != 1) epi_res = 2; }
  const size_t smem_res = 1008 + (size_t)P.k_blocks * kBTileBytes + (size_t)kGemmStagesRes * kATileBytes + 2 * kStageOutBytes +
                          128 + 2 * kGemmBN * sizeof(float);
  const bool resident = w_res_knob && smem_res <= 232448 && P.n_blocks <= num_sms() && P.m_blocks >= 4 * (num_sms() / P.n_blocks);
  if (resident) {
    static size_t attr_r = 0;
    if (smem_res > attr_r) {
      VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_kernel<true, 1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res));
      VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_kernel<true, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res));
      VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_kernel<true, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_res));
      attr_r = smem_res;
    }
    const int groups = num_sms() / P.n_blocks;
    if (epi_res == 1) gemm_geglu_kernel<true, 1, 1><<<groups * P.n_blocks, 256, smem_res, (cudaStream_t)stream>>>(ma, mw, mo, P);
    else if (gelu_knob) gemm_geglu_kernel<true, 2, 1><<<groups * P.n_blocks, kGemmThreads, smem_res, (cudaStream_t)stream>>>(ma, mw, mo, P);
    else gemm_geglu_kernel<true, 2, 0><<<groups * P.n_blocks, kGemmThreads, smem_res, (cudaStream_t)stream>>>(ma, mw, mo, P);
    return check_cuda(cudaGetLastError(), "gemm_geglu_kernel<resident W> launch");
  }
  const size_t smem = 1008 + (size_t)kGemmStages * kStageBytes + 2 * kStageOutBytes + kTail;
  static_assert(1008 + (size_t)kGemmStages * kStageBytes + 2 * kStageOutBytes + kTail <= 232448, "streaming kernel shared memory");
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_kernel<false, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    VF_CUDA_TRY(cudaFuncSetAttribute(gemm_geglu_kernel<false, 2, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  long long grid = (long long)P.m_blocks * P.n_blocks;
  if (grid > num_sms()) grid = num_sms();
  if (gelu_knob) gemm_geglu_kernel<false, 2, 1><<<(int)grid, kGemmThreads, smem, (cudaStream_t)stream>>>(ma, mw, mo, P);
  else gemm_geglu_kernel<false, 2, 0><<<(int)grid, kGemmThreads, smem, (cudaStream_t)stream>>>(ma, mw, mo, P);
  return check_cuda(cudaGetLastError(), "gemm_geglu_kernel launch");
}
