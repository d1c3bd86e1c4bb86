// Projection GEMM of the 64x64 transformer level on tcgen05, with everything that surrounds it in the reference folded in:
//
//   out[r, 0:n] = LN(x[r]) . W^T + bias[sample(r)] + residual[r]          (each of LN, bias, residual optional)
//
// Replaces, for k <= 320 (the 64x64 level, where 4/5 of the projection time of a step is spent):
//   * norm1 + to_q / to_k / to_v      ldm/modules/attention.py:239 (`self.attn1(self.norm1(x))`) + :172-174 / pnp_utils.py:106
//   * to_out + bias + attn2 row + x   attention.py:176,221 / pnp_utils.py:287 and the adds of attention.py:239-241
//   * proj_in (1x1 conv)              attention.py:261-265,279
//   * proj_out (1x1 conv) + x_in      attention.py:270-274,287-288
// which the first version of this path ran as LayerNorm kernel + cuBLAS / cuBLASLt GEMMs.
//
// These launches are HBM-bound (k = 320: 0.24 TFLOP against 1.0 GB for QKV), so the design goal is to move every operand
// through HBM exactly once and keep enough bytes in flight per SM:
//   * W-resident: a CTA owns ONE 160-column slice of the output (100 KB of W for all of k, loaded once) and walks the row
//     blocks of its group; A streams through a 5-stage ring of 128 x 64 k-blocks (80 KB in flight per SM); the n / 160 CTAs
//     of a group read the same A tiles at the same time, so A crosses HBM once and L2 n / 160 times.
//   * LayerNorm without a pass over x: LN(x) . W^T = rstd * (x . (W o gamma)^T - mean * colsum(W o gamma)) + beta . W^T.  The MMA runs
//     on the RAW rows; four "stats" warps read the same shared-memory k-blocks the tensor core reads (thread = row) and
//     accumulate sum and sum of squares, and the epilogue applies mean / rstd per row.  The normalised tensor is never
//     written or read: 2 of the 6 row-units of LN + QKV disappear, and x-hat is never rounded to bf16.
//   * the residual add is three more k-blocks of the SAME pipeline: the residual tile travels through the A ring as 128 x 64
//     boxes and is accumulated with MMAs against a 64 x 64 bf16 identity kept in shared memory (acc[:, j] += r[:, j] * 1.0 is
//     exact in the fp32 accumulator), so it needs no second ring, no generic-proxy reads and no epilogue work, and is
//     prefetched exactly as deep as A.  (The first version staged it next to the output: its two-deep ring exposed the L2
//     latency of every box.)
//   * output through one 16 KB staging buffer per epilogue group, in steps of 64 / 64 / 32 columns, drained by TMA stores.  No row-per-thread global access anywhere
//     (the failure mode of the first GEMM+GEGLU epilogue, profiles/r1_geglu_gemm_ncu.txt).
//   * two accumulator sets of 160 TMEM columns: the epilogue of tile i runs under the MMAs of tile i + 1.
//
//   * two epilogue warp groups, one per accumulator set (tiles alternate between them), each with its own half of the
//     staging ring and its own bulk-store groups: a tile's epilogue is a chain of
//     long-latency steps (tcgen05.ld, fence.proxy.async, named barrier, store issue) and one group alone set the pace.
//   * the next tile's A k-blocks and residual boxes are pulled into L2 one tile ahead (cp.async.bulk.prefetch.tensor), so
//     the loads that fill shared memory are L2 hits and the rings hold bandwidth x L2 latency, not x HBM latency.
//
// Warp roles (512 threads, one CTA per SM): 0 TMA producer (W once, then A and residual boxes), 1 TMEM + MMA issuer, 2-3 idle,
// 4-7 / 8-11 epilogue groups 0 / 1 (TMEM lane quarter = warp % 4), 12-15 row statistics (LayerNorm form only; LayerNorm and
// residual are mutually exclusive: the residual inside the accumulator would be scaled by rstd).
#include "vf_common.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cstdlib>

namespace vf {

using namespace sm100;

constexpr int kP3Threads = 512;
constexpr int kP3BM = 128;            // rows per tile
constexpr int kP3BN = 160;            // output columns per CTA (one MMA: M128 N160)
constexpr int kP3BK = 64;             // k-block: one 128-byte swizzled row
constexpr int kP3Stages = 5;          // A ring depth (a whole k = 320 tile)
constexpr int kP3OutBufs = 4;         // residual / output staging ring
constexpr int kP3MaxKBlocks = 5;      // k <= 320
constexpr uint32_t kP3ATile = kP3BM * kP3BK * 2;          // 16 KB
constexpr uint32_t kP3WTile = kP3BN * kP3BK * 2;          // 20 KB
constexpr uint32_t kP3OutBytes = kP3BM * 32 * 2;          // 8 KB: 128 rows x 32 columns
constexpr int kP3Steps = kP3BN / 32;                      // column steps of the epilogue per tile

struct Proj3Params {
  const float* bias;               // (bias_rows, n) fp32 or null
  const float* colsum;             // (n) fp32: sum over k of the (gamma-folded, bf16) weight row; LayerNorm form only
  const float2* stats_in;          // kLN == 2: (rows, stats_parts) partial {sum, sum of squares} of the input rows
  float2* stats_out;               // kEmit: (rows, n_blocks) partial {sum, sum of squares} of the bf16 OUTPUT rows
  int stats_parts;
  long long rows;
  long long rows_per_bias;         // 0: one bias row for all; else rows sharing one bias row (multiple of 128)
  int n, k;
  int m_blocks, n_blocks, k_blocks;
  int has_residual;
  int prefetch;                    // bit 0: next tile's A into L2, bit 1: the group's next residual boxes into L2
  float eps, inv_k;
};

struct __align__(8) Proj3Barriers {
  uint64_t full[kP3Stages], empty[kP3Stages];
  uint64_t acc_full[2], acc_empty[2];
  uint64_t stats_full[3];
  uint64_t w_full;
  uint32_t tmem_base;
};
static_assert(sizeof(Proj3Barriers) <= 256, "barrier block");

__device__ __forceinline__ void p3_tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      :: "r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void p3_tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];"
               :: "l"(reinterpret_cast<uint64_t>(m)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void p3_bar_sync(int id, int threads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// The CTA owns n-block blockIdx.x % n_blocks and walks row blocks g, g + groups, ... (groups = gridDim.x / n_blocks).
struct P3Walk {
  int cur, end, step, nb;
  __device__ P3Walk(const Proj3Params& P) {
    nb = blockIdx.x % P.n_blocks;
    cur = blockIdx.x / P.n_blocks;
    step = gridDim.x / P.n_blocks;
    end = P.m_blocks;
  }
  __device__ bool valid() const { return cur < end; }
  __device__ void next() { cur += step; }
};

// kLN: 0 no normalisation, 1 row statistics computed here (stats warps), 2 row statistics handed over by the kernel that
// produced x (its kEmit epilogue).  kEmit: the epilogue also writes {sum, sum of squares} of every bf16 output row of its
// 160-column slice, so that a LayerNorm over the output (the next projection's kLN == 2) needs no pass of its own.
template <int kLN, bool kEmit>
__global__ void __launch_bounds__(kP3Threads, 1)
proj3_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
             const __grid_constant__ CUtensorMap map_r, const __grid_constant__ CUtensorMap map_o,
             const __grid_constant__ CUtensorMap map_o32, const Proj3Params P) {
  // dynamic shared memory: [pad to 1024] | W (k_blocks x 20 KB) | A ring (5 x 16 KB) | staging (4 x 8 KB) | identity (8 KB) |
  //                        barriers | bias[2][160] | colsum[160] | stats[3][128]
  extern __shared__ unsigned char p3_smem[];
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t dyn_base = smem_u32(p3_smem);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  unsigned char* w_res = p3_smem + (tile_base - dyn_base);
  unsigned char* ring = w_res + (size_t)P.k_blocks * kP3WTile;
  unsigned char* out_stage = ring + (size_t)kP3Stages * kP3ATile;
  unsigned char* ident = out_stage + (size_t)kP3OutBufs * kP3OutBytes;      // I_64, K-major, 128B swizzle (residual MMAs)
  unsigned char* tail = ident + 64 * 128;
  Proj3Barriers& bars = *reinterpret_cast<Proj3Barriers*>(tail);
  float (*s_bias)[kP3BN] = reinterpret_cast<float (*)[kP3BN]>(tail + 256);
  float* s_cs = reinterpret_cast<float*>(tail + 256 + 2 * kP3BN * sizeof(float));
  float2 (*s_stats)[kP3BM] = reinterpret_cast<float2 (*)[kP3BM]>(tail + 256 + 3 * kP3BN * sizeof(float));

  if (threadIdx.x == 0) {
    for (int s = 0; s < kP3Stages; ++s) {
      mbar_init(&bars.full[s], 1);
      mbar_init(&bars.empty[s], kLN == 1 ? 5 : 1);               // the MMA commit (+ one arrival per stats warp)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars.acc_full[a], 1);
      mbar_init(&bars.acc_empty[a], 4);                     // one arrival per epilogue warp
    }
    for (int b = 0; b < 3; ++b) mbar_init(&bars.stats_full[b], 4);
    mbar_init(&bars.w_full, 1);
    fence_barrier_init();
  }
  if (P.has_residual) {
    for (int i = threadIdx.x; i < 64 * 128 / 16; i += kP3Threads) reinterpret_cast<uint4*>(ident)[i] = make_uint4(0u, 0u, 0u, 0u);
    __syncthreads();
    if (threadIdx.x < 64) {          // element (n, n): row n, 16-byte chunk (n / 8) ^ (n % 8), position n % 8
      const int nn = threadIdx.x;
      *reinterpret_cast<__nv_bfloat16*>(ident + nn * 128 + (((nn >> 3) ^ (nn & 7)) << 4) + (nn & 7) * 2) = __float2bfloat16_rn(1.0f);
    }
    fence_proxy_async();             // generic-proxy writes -> visible to the tensor core's shared-memory reads
  }
  if (warp == 1) tmem_alloc<512>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp == 0) {
    // =========================== TMA producer: W once, then the A k-blocks of every tile ==================
    if (lane == 0) {
      tma_prefetch_desc(&map_a);
      tma_prefetch_desc(&map_w);
      P3Walk tw(P);
      if (tw.valid()) {
        mbar_arrive_expect_tx(&bars.w_full, (uint32_t)P.k_blocks * kP3WTile);
        for (int kb = 0; kb < P.k_blocks; ++kb)
          p3_tma_load_2d(w_res + (size_t)kb * kP3WTile, &map_w, &bars.w_full, kb * kP3BK, tw.nb * kP3BN);
      }
      uint32_t it = 0;
      for (; tw.valid(); tw.next()) {
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kP3Stages;
          mbar_wait(&bars.empty[s], ((it / kP3Stages) & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.full[s], kP3ATile);
          p3_tma_load_2d(ring + (size_t)s * kP3ATile, &map_a, &bars.full[s], kb * kP3BK, tw.cur * kP3BM);
          if ((P.prefetch & 1) && tw.cur + tw.step < tw.end) p3_tma_prefetch_2d(&map_a, kb * kP3BK, (tw.cur + tw.step) * kP3BM);
        }
        if (P.has_residual) {
          // residual columns [0,64), [64,128) and [96,160) of this CTA's slice (the last box overlaps: its upper half is used)
          for (int r = 0; r < 3; ++r, ++it) {
            const int s = it % kP3Stages;
            const int col = tw.nb * kP3BN + (r == 2 ? 96 : r * 64);
            mbar_wait(&bars.empty[s], ((it / kP3Stages) & 1) ^ 1);
            mbar_arrive_expect_tx(&bars.full[s], kP3ATile);
            p3_tma_load_2d(ring + (size_t)s * kP3ATile, &map_r, &bars.full[s], col, tw.cur * kP3BM);
            if ((P.prefetch & 2) && tw.cur + tw.step < tw.end) p3_tma_prefetch_2d(&map_r, col, (tw.cur + tw.step) * kP3BM);
          }
        }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer ================================================================
    if (lane == 0) {
      const uint32_t idesc = make_idesc_bf16(kP3BM, kP3BN, false);
      const uint32_t idesc_r64 = make_idesc_bf16(kP3BM, 64, false);
      const uint32_t idesc_r32 = make_idesc_bf16(kP3BM, 32, false);
      uint32_t it = 0, ti = 0;
      P3Walk tw(P);
      if (tw.valid()) {
        mbar_wait(&bars.w_full, 0);
        tc_fence_after();
      }
      for (; tw.valid(); tw.next(), ++ti) {
        const uint32_t ab = ti & 1;
        mbar_wait(&bars.acc_empty[ab], ((ti >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t acc = tmem + ab * 256;
        for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
          const int s = it % kP3Stages;
          mbar_wait(&bars.full[s], (it / kP3Stages) & 1);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(ring + (size_t)s * kP3ATile);
          const uint32_t b_addr = smem_u32(w_res + (size_t)kb * kP3WTile);
#pragma unroll
          for (int ks = 0; ks < kP3BK / 16; ++ks) {
            const uint64_t da = make_smem_desc_sw128(a_addr + ks * 32, 16, 1024);
            const uint64_t db = make_smem_desc_sw128(b_addr + ks * 32, 16, 1024);
            mma_ss(acc, da, db, idesc, (kb > 0) || (ks > 0));
          }
          tc_commit(&bars.empty[s]);
        }
        if (P.has_residual) {
          const uint32_t i_addr = smem_u32(ident);
          for (int r = 0; r < 3; ++r, ++it) {
            const int s = it % kP3Stages;
            mbar_wait(&bars.full[s], (it / kP3Stages) & 1);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(ring + (size_t)s * kP3ATile);
            if (r < 2) {
              // acc[:, 64 r + j] += box[:, j] for j < 64
#pragma unroll
              for (int ks = 0; ks < 4; ++ks)
                mma_ss(acc + r * 64, make_smem_desc_sw128(a_addr + ks * 32, 16, 1024), make_smem_desc_sw128(i_addr + ks * 32, 16, 1024),
                       idesc_r64, true);
            } else {
              // acc[:, 128 + j] += box[:, 32 + j] for j < 32: k-slices 2, 3 of the box against rows 32..63 of the identity
#pragma unroll
              for (int ks = 2; ks < 4; ++ks)
                mma_ss(acc + 128, make_smem_desc_sw128(a_addr + ks * 32, 16, 1024),
                       make_smem_desc_sw128(i_addr + 32 * 128 + ks * 32, 16, 1024), idesc_r32, true);
            }
            tc_commit(&bars.empty[s]);
          }
        }
        tc_commit(&bars.acc_full[ab]);
      }
    }
  } else if (warp >= 4 && warp < 12) {
    // =========================== epilogue ==================================================================
    const int quarter = warp & 3;
    const uint32_t grp = (uint32_t)(warp - 4) >> 2;         // this group takes the tiles of accumulator set `grp`
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int et = (threadIdx.x - 128) & 127;               // 0..127 within the group
    const int bar0 = 1 + 3 * (int)grp;                      // named barriers of the group
    const int trow = quarter * 32 + lane;                   // row of the tile this thread owns
    const uint32_t sw = (uint32_t)(trow >> 1) & 3u;         // 64B swizzle of the staging rows
    const bool issuer = et == 0;
    const bool has_bias = P.bias != nullptr;
    if (issuer) { tma_prefetch_desc(&map_o); tma_prefetch_desc(&map_o32); }
    uint32_t ti = 0;
    P3Walk tw(P);
    if (kLN)
      for (int i = et; i < kP3BN; i += 128) s_cs[i] = P.colsum[tw.nb * kP3BN + i];
    for (; tw.valid(); tw.next(), ++ti) {
      const uint32_t ab = ti & 1;
      if (ab != grp) continue;
      {
        const long long brow = P.rows_per_bias ? ((long long)tw.cur * kP3BM) / P.rows_per_bias : 0;
        for (int i = et; i < kP3BN; i += 128)
          s_bias[ab][i] = P.bias ? P.bias[brow * P.n + tw.nb * kP3BN + i] : 0.0f;
        p3_bar_sync(bar0, 128);
      }
      float mu_rs = 0.0f, rs = 1.0f;                        // -mean * rstd, rstd
      const long long grow = (long long)tw.cur * kP3BM + trow;      // this thread's row of the tensor
      if (kLN == 1) {
        mbar_wait(&bars.stats_full[ti % 3], (ti / 3) & 1);
        const float2 st = s_stats[ti % 3][trow];
        rs = st.y;
        mu_rs = -st.x * st.y;
      } else if (kLN == 2) {
        float sx = 0.0f, sq = 0.0f;
        if (grow < P.rows)
          for (int j = 0; j < P.stats_parts; ++j) {
            const float2 pp = P.stats_in[grow * P.stats_parts + j];
            sx += pp.x;
            sq += pp.y;
          }
        const float mean = sx * P.inv_k;
        rs = rsqrtf(fmaxf(sq * P.inv_k - mean * mean, 0.0f) + P.eps);
        mu_rs = -mean * rs;
      }
      float2 e_sum = make_float2(0.0f, 0.0f), e_sq = make_float2(0.0f, 0.0f);
      mbar_wait(&bars.acc_full[ab], (ti >> 1) & 1);
      tc_fence_after();
      const uint32_t acc = tmem + ab * 256 + lane_off;
      // three column steps per tile: 64, 64 and 32 columns.  The staging rows are as wide as the step (128-byte rows with
      // the 128B swizzle, 64-byte rows with the 64B swizzle for the last one): the TMA store engine takes a fixed time per
      // box ROW, and 64-byte rows capped the first version of this epilogue at ~10 B/clk/SM (3900 cycles per 40 KB tile for
      // every shape, profiles/r2_proj_v2_ncu.txt).  One 16 KB buffer per group; the drain of step i overlaps the TMEM loads
      // and the arithmetic of step i + 1.
      unsigned char* obuf = out_stage + (size_t)grp * (2 * kP3OutBytes);
#pragma unroll 1
      for (int st = 0; st < 3; ++st) {
        const int c0 = st * 64;
        const int halves = st < 2 ? 2 : 1;
        uint32_t pk[32];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < halves) {
            uint32_t v[32];
            tmem_ld_x32(acc + c0 + h * 32, v);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              // four columns per iteration: one 16-byte broadcast read per constant vector (the shared-memory pipe, which
              // also feeds the tensor core's operands, is the busiest unit of this kernel)
              const int i = j * 4;
              float2 a0 = make_float2(__uint_as_float(v[i]), __uint_as_float(v[i + 1]));
              float2 a1 = make_float2(__uint_as_float(v[i + 2]), __uint_as_float(v[i + 3]));
              float4 bb = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
              if (kLN || has_bias) bb = *reinterpret_cast<const float4*>(&s_bias[ab][c0 + h * 32 + i]);
              if (kLN) {
                const float4 cs = *reinterpret_cast<const float4*>(&s_cs[c0 + h * 32 + i]);
                const float2 m2 = make_float2(mu_rs, mu_rs), r2 = make_float2(rs, rs);
                a0 = __ffma2_rn(r2, a0, __ffma2_rn(m2, make_float2(cs.x, cs.y), make_float2(bb.x, bb.y)));      // rstd acc + (bias - mean rstd colsum)
                a1 = __ffma2_rn(r2, a1, __ffma2_rn(m2, make_float2(cs.z, cs.w), make_float2(bb.z, bb.w)));
              } else if (has_bias) {
                a0 = __fadd2_rn(a0, make_float2(bb.x, bb.y));
                a1 = __fadd2_rn(a1, make_float2(bb.z, bb.w));
              }
              const uint32_t p0 = pack_bf16(a0.x, a0.y), p1 = pack_bf16(a1.x, a1.y);
              pk[h * 16 + j * 2] = p0;
              pk[h * 16 + j * 2 + 1] = p1;
              if (kEmit) {                                    // statistics of the ROUNDED values: what the consumer will read
                const float2 r0 = make_float2(bf16lo(p0), bf16hi(p0)), r1 = make_float2(bf16lo(p1), bf16hi(p1));
                e_sum = __fadd2_rn(e_sum, __fadd2_rn(r0, r1));
                e_sq = __ffma2_rn(r0, r0, __ffma2_rn(r1, r1, e_sq));
              }
            }
          }
        }
        if (issuer) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");     // the group's previous store has drained the buffer
        p3_bar_sync(bar0 + 1, 128);
        if (st < 2) {
          const uint32_t sw8 = (uint32_t)trow & 7u;           // 128B swizzle: 16-byte chunk q of row r sits at chunk q ^ (r & 7)
#pragma unroll
          for (int q = 0; q < 8; ++q)
            *reinterpret_cast<uint4*>(obuf + trow * 128 + ((q ^ sw8) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(obuf + trow * 64 + ((q ^ sw) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        }
        fence_proxy_async();
        p3_bar_sync(bar0 + 2, 128);
        if (issuer) {
          const CUtensorMap* mo = st < 2 ? &map_o : &map_o32;
          asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];"
                       :: "l"(reinterpret_cast<uint64_t>(mo)), "r"(tw.nb * kP3BN + c0), "r"(tw.cur * kP3BM), "r"(smem_u32(obuf))
                       : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      if (kEmit && grow < P.rows) P.stats_out[grow * P.n_blocks + tw.nb] = make_float2(e_sum.x + e_sum.y, e_sq.x + e_sq.y);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.acc_empty[ab]);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");      // stores complete before exit
  } else if (kLN == 1 && warp >= 12) {
    // =========================== row statistics (LayerNorm form) ===========================================
    const int row = (warp - 12) * 32 + lane;
    const uint32_t swz = (uint32_t)row & 7u;                // 128B swizzle of the A k-block rows
    uint32_t it = 0, ti = 0;
    for (P3Walk tw(P); tw.valid(); tw.next(), ++ti) {
      float2 sum = make_float2(0.0f, 0.0f), sq = make_float2(0.0f, 0.0f);
      for (int kb = 0; kb < P.k_blocks; ++kb, ++it) {
        const int s = it % kP3Stages;
        mbar_wait(&bars.full[s], (it / kP3Stages) & 1);
        const unsigned char* rp = ring + (size_t)s * kP3ATile + row * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const uint4 u4 = *reinterpret_cast<const uint4*>(rp + ((q ^ swz) << 4));
          const uint32_t uu[4] = {u4.x, u4.y, u4.z, u4.w};
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 x2 = make_float2(bf16lo(uu[j]), bf16hi(uu[j]));
            sum = __fadd2_rn(sum, x2);
            sq = __ffma2_rn(x2, x2, sq);
          }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.empty[s]);
      }
      const float mean = (sum.x + sum.y) * P.inv_k;
      const float var = fmaxf((sq.x + sq.y) * P.inv_k - mean * mean, 0.0f);
      s_stats[ti % 3][row] = make_float2(mean, rsqrtf(var + P.eps));
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.stats_full[ti % 3]);
    }
  }

  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

typedef CUresult (*EncodeTiledFn3)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn3 p3_encode_fn() {
  static EncodeTiledFn3 fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn3>(p);
  return fn;
}

// (rows, cols) bf16 row-major with row stride ld elements -> 2-D map {cols, rows}, box {box_c, box_r}
static int p3_make_map(CUtensorMap* m, const void* base, long long rows, int cols, long long ld, int box_c, int box_r,
                       CUtensorMapSwizzle swz, CUtensorMapL2promotion promo) {
  EncodeTiledFn3 enc = p3_encode_fn();
  if (!enc) return fail("vf_linear_proj: cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_c, (cuuint32_t)box_r};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("vf_linear_proj: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace vf

extern "C" int vf_linear_proj_supported(long long rows, int k, int n) {
  using namespace vf;
  return rows > 0 && k >= kP3BK && k % kP3BK == 0 && k <= kP3MaxKBlocks * kP3BK && n > 0 && n % kP3BN == 0 &&
         n / kP3BN <= 148;
}

extern "C" int vf_linear_proj(const void* x, const void* w, const float* bias, long long rows_per_bias, const void* residual,
                              const float* ln_colsum, float ln_eps, const float* ln_stats_in, int ln_stats_parts,
                              float* stats_out, void* out, long long rows, int k, int n, long long ld_x,
                              long long ld_res, long long ld_out, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !w || !out) return fail("vf_linear_proj: null pointer");
  if (dtype != VF_BF16) return fail("vf_linear_proj: bf16 only (fp32 runs go through the library GEMM)");
  if (!vf_linear_proj_supported(rows, k, n))
    return fail("vf_linear_proj: bad shape rows=%lld k=%d n=%d (k a multiple of %d up to %d, n a multiple of %d)", rows, k, n,
                kP3BK, kP3MaxKBlocks * kP3BK, kP3BN);
  if (ld_x < k || ld_x % 8 || ld_out < n || ld_out % 8 || (residual && (ld_res < n || ld_res % 8)))
    return fail("vf_linear_proj: bad row stride (ld_x %lld, ld_res %lld, ld_out %lld; multiples of 8 elements)", ld_x, ld_res, ld_out);
  if (rows_per_bias < 0 || (rows_per_bias > 0 && (!bias || rows_per_bias % kP3BM || rows % rows_per_bias)))
    return fail("vf_linear_proj: rows_per_bias %lld must be a multiple of %d that divides rows (and needs a bias)", rows_per_bias, kP3BM);
  if (out == residual || out == x) return fail("vf_linear_proj: out must not alias x or residual");
  if (ln_stats_in && (!ln_colsum || ln_stats_parts < 1 || ln_stats_parts > 16))
    return fail("vf_linear_proj: ln_stats_in needs the LayerNorm form and 1..16 partials per row (got %d)", ln_stats_parts);
  if (stats_out && ln_colsum) return fail("vf_linear_proj: stats_out is not available in the LayerNorm form");
  if ((reinterpret_cast<uintptr_t>(ln_stats_in) | reinterpret_cast<uintptr_t>(stats_out)) & 7)
    return fail("vf_linear_proj: row statistics must be 8-byte aligned");
  if (residual && ln_colsum) return fail("vf_linear_proj: the LayerNorm form takes no residual (it would be scaled by rstd inside the accumulator)");
  const void* ptrs[4] = {x, w, out, residual};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail("vf_linear_proj: pointers must be 16-byte aligned");

  Proj3Params P;
  P.bias = bias;
  P.colsum = ln_colsum;
  P.stats_in = reinterpret_cast<const float2*>(ln_stats_in);
  P.stats_out = reinterpret_cast<float2*>(stats_out);
  P.stats_parts = ln_stats_parts;
  P.rows = rows;
  P.rows_per_bias = rows_per_bias;
  P.n = n; P.k = k;
  P.m_blocks = (int)((rows + kP3BM - 1) / kP3BM);
  P.n_blocks = n / kP3BN;
  P.k_blocks = k / kP3BK;
  P.has_residual = residual ? 1 : 0;
  static int pf_knob = -1;         // VF_PROJ_PREFETCH: bit 0 next A tile, bit 1 next residual boxes (default both)
  if (pf_knob < 0) { const char* e = getenv("VF_PROJ_PREFETCH"); pf_knob = e ? atoi(e) : 3; }
  P.prefetch = pf_knob;
  P.eps = ln_eps;
  P.inv_k = 1.0f / (float)k;

  CUtensorMap ma, mw, mr, mo;
  if (int rc = p3_make_map(&ma, x, rows, k, ld_x, kP3BK, kP3BM, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return rc;
  if (int rc = p3_make_map(&mw, w, n, k, k, kP3BK, kP3BN, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return rc;
  CUtensorMap mo32;
  if (int rc = p3_make_map(&mo, out, rows, n, ld_out, 64, kP3BM, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE)) return rc;
  if (int rc = p3_make_map(&mo32, out, rows, n, ld_out, 32, kP3BM, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE)) return rc;
  if (residual) {
    if (int rc = p3_make_map(&mr, residual, rows, n, ld_res, kP3BK, kP3BM, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B)) return rc;
  } else {
    mr = ma;
  }

  const size_t smem = 1008 + (size_t)P.k_blocks * kP3WTile + (size_t)kP3Stages * kP3ATile + (size_t)kP3OutBufs * kP3OutBytes +
                      64 * 128 + 256 + 3 * kP3BN * sizeof(float) + 3 * kP3BM * sizeof(float2);
  if (smem > 232448) return fail("vf_linear_proj: shared-memory plan %zu exceeds the 227 KB of an SM", smem);
  int dev = 0;
  VF_CUDA_TRY(cudaGetDevice(&dev));
  // variant: 0 plain, 1 plain + emitted statistics, 2 LayerNorm with in-kernel statistics, 3 LayerNorm with handed-over statistics
  const int variant = ln_colsum ? (ln_stats_in ? 3 : 2) : (stats_out ? 1 : 0);
  typedef void (*KernelFn)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const Proj3Params);
  static const KernelFn kernels[4] = {proj3_kernel<0, false>, proj3_kernel<0, true>, proj3_kernel<1, false>, proj3_kernel<2, false>};
  static size_t attr_dev[64][4] = {};
  if (dev < 64 && smem > attr_dev[dev][variant]) {
    VF_CUDA_TRY(cudaFuncSetAttribute(kernels[variant], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_dev[dev][variant] = smem;
  }
  int groups = num_sms() / P.n_blocks;
  if (groups > P.m_blocks) groups = P.m_blocks;
  if (groups < 1) groups = 1;
  const int grid = groups * P.n_blocks;
  kernels[variant]<<<grid, kP3Threads, smem, (cudaStream_t)stream>>>(ma, mw, mr, mo, mo32, P);
  return check_cuda(cudaGetLastError(), "proj3_kernel launch");
}
