// Fused attention, "round-robin" arrangement: ONE CTA per SM works on THREE 128-row query tiles of one (batch, head)
// and its twelve softmax warps take turns on the exponential unit.
//
// Why.  At d_head = 40 the kernel of vf_attn_tc.cu is bound by MUFU.EX2 (one exponential per 160 MMA flops, 16 per clock
// and SM) and sits at ~0.75 of that ceiling whatever the knob: its three CTAs per SM are independent,
// so whether the three softmax warps that share a scheduler are in their exponential sections at the same time -- the
// XU then serves them a third each while nobody does the latency-bound part of a tile -- or all outside it -- the XU
// idles -- is left to chance.  Here the three warps of a scheduler belong to the same CTA (tiles A, B, C of one CTA,
// same TMEM lane quarter) and pass a token around with named barriers: A's exponentials, then B's, then C's, then A's
// of the next key tile.  The exponential section of a warp runs alone on its scheduler's XU (MUFU-paced, 64 MUFU = 512
// cycles) while the other two do their S load / row max / P store / barrier round trips, which need issue slots only.
//
// Otherwise the data path is that of vf_attn_tc.cu (split-P): S = Q K^T (SS MMA) into TMEM, softmax in registers (one
// thread per query row), P (bf16) to TMEM, O += P V (TS MMA), lazy rescale of O, K/V through a TMA ring -- shared by the
// three tiles, so a CTA pulls each key/value tile once for 384 query rows instead of once per 128.
//   TMEM  512 columns: tile t at 160 t: S fp32 [0,64) | O fp32 [64,128) | P bf16x2 [128,160)
//   warps 0-11 softmax (tile = warp / 4, lane quarter = warp % 4), 12 TMA + TMEM allocator, 13-15 MMA issuers (one per tile)
//   d_head <= 64 only (the UNet's 40); everything else stays on vf_attn_tc.cu.
#include "vf_attn.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cstdlib>

namespace vf {

using namespace sm100;

int attn_make_map(CUtensorMap* m, const void* base, int batch, int heads, int n, int d, long long ld, int box_rows);

constexpr int kPpTiles = 3;
constexpr int kPpThreads = 512;
constexpr int kPpBM = 128;
constexpr int kPpBN = 64;
constexpr int kPpStages = 4;
constexpr int kPpTmaWarp = 12;      // TMA producer and TMEM allocator; warps 13, 14, 15: MMA issuers of tiles 0, 1, 2
constexpr int kPpTileCols = 160;
constexpr float kPpRescaleThreshold = 8.0f;   // log2 units
#ifndef VF_PP_SIGNAL_AT
#define VF_PP_SIGNAL_AT 40
#endif
constexpr int kPpSignalAt = VF_PP_SIGNAL_AT;  // exponentials of a tile issued before the turn is handed on (even, < 64)

struct AttnPpParams {
  __nv_bfloat16* o;
  long long ld_o;
  int heads, n_q, n_kv, n_kv2, d, d_pad;
  float scale_log2;
  int order;            // softmax warps of a scheduler allowed in their exponential sections at a time (1, 2); 0: no ordering
  int one;              // 1 (trip count of the single-pass loop that fences the consumers of the exponentials off)
};

struct __align__(8) PpBarriers {
  uint64_t q_full;
  uint64_t k_full[kPpStages], k_empty[kPpStages], v_full[kPpStages], v_empty[kPpStages];
  uint64_t s_full[kPpTiles], s_free[kPpTiles], p_full[kPpTiles], p_empty[kPpTiles], o_done[kPpTiles];
  uint32_t tmem_base;
  uint32_t turn[kPpTiles][4];   // exponential sections finished by softmax warp (tile, quarter)
  uint32_t zero;                // 0
};

template <int kRegs> __device__ __forceinline__ void pp_reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void pp_reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kRegs)); }
__device__ __forceinline__ void pp_bar_sync(int id) { asm volatile("bar.sync %0, 64;" :: "r"(id) : "memory"); }
__device__ __forceinline__ void pp_bar_arrive(int id) { asm volatile("bar.arrive %0, 64;" :: "r"(id) : "memory"); }
// volatile: the 64 exponentials stay one uninterrupted stream in program order
__device__ __forceinline__ float pp_ex2(float x) {
  float y;
  asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kPpThreads, 1)
attn_pp_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_k2,
               const __grid_constant__ CUtensorMap map_v2, const AttnPpParams P) {
  constexpr int BN = kPpBN;
  extern __shared__ unsigned char smem_dyn[];
  __shared__ PpBarriers bars;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_group = blockIdx.x;                      // 3 consecutive 128-row query tiles
  const int bh = blockIdx.y;
  const int b = bh / P.heads, h = bh - b * P.heads;

  const uint32_t dyn_base = smem_u32(smem_dyn);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  unsigned char* tiles = smem_dyn + (tile_base - dyn_base);
  constexpr uint32_t q_bytes = kPpBM * 128;            // one 128 x 64 bf16 tile, 128B-swizzled rows
  constexpr uint32_t kv_bytes = BN * 128;
  unsigned char* sQ = tiles;                           // 3 tiles
  unsigned char* sK = sQ + kPpTiles * q_bytes;         // kPpStages stages
  unsigned char* sV = sK + kPpStages * kv_bytes;

  const int t1 = (P.n_kv + BN - 1) / BN;
  const int t2 = (P.n_kv2 + BN - 1) / BN;
  const int n_tiles = t1 + t2;

  if (threadIdx.x == 0) {
    for (int i = 0; i < kPpTiles * 4; ++i) (&bars.turn[0][0])[i] = 0u;
    bars.zero = 0u;
    mbar_init(&bars.q_full, 1);
    for (int s = 0; s < kPpStages; ++s) {
      mbar_init(&bars.k_full[s], 1);
      mbar_init(&bars.k_empty[s], kPpTiles);          // released by the three tiles' issuers
      mbar_init(&bars.v_full[s], 1);
      mbar_init(&bars.v_empty[s], kPpTiles);
    }
    for (int t = 0; t < kPpTiles; ++t) {
      mbar_init(&bars.s_full[t], 1);
      mbar_init(&bars.s_free[t], 4);
      mbar_init(&bars.p_full[t], 4);
      mbar_init(&bars.p_empty[t], 1);
      mbar_init(&bars.o_done[t], 1);
    }
    fence_barrier_init();
  }
  if (warp == kPpTmaWarp) tmem_alloc<512>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;

  if (warp >= 12) {
    pp_reg_dec<40>();
    if (warp == kPpTmaWarp) {
      // =========================== TMA producer ==================================================
      if (lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_arrive_expect_tx(&bars.q_full, kPpTiles * q_bytes);
        for (int t = 0; t < kPpTiles; ++t)               // rows beyond n_q are zero-filled
          tma_load_4d(sQ + t * q_bytes, &map_q, &bars.q_full, 0, h, (q_group * kPpTiles + t) * kPpBM, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int st = j % kPpStages;
          const uint32_t use = (uint32_t)(j / kPpStages);
          const bool seg2 = j >= t1;
          const int row0 = (seg2 ? j - t1 : j) * BN;
          const CUtensorMap* mk = seg2 ? &map_k2 : &map_k;
          const CUtensorMap* mv = seg2 ? &map_v2 : &map_v;
          mbar_wait(&bars.k_empty[st], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.k_full[st], kv_bytes);
          tma_load_4d(sK + st * kv_bytes, mk, &bars.k_full[st], 0, h, row0, b);
          mbar_wait(&bars.v_empty[st], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.v_full[st], kv_bytes);
          tma_load_4d(sV + st * kv_bytes, mv, &bars.v_full[st], 0, h, row0, b);
        }
      }
    } else {
      // =========================== MMA issuers: warp 13 + t serves tile t (both chains) ===============
      // One issuer for all three tiles (first version) walked them in order and every barrier probe is a 300-cycle round
      // trip under load: S_t(j+1) then arrived late and 28 % of all warp-stall samples sat on s_full.
      const int t = warp - 13;
      const uint32_t idesc_qk = make_idesc_bf16(kPpBM, BN, false);
      const uint32_t idesc_pv = make_idesc_bf16(kPpBM, P.d_pad, true);
      const int k_steps = P.d_pad / 16;
      const uint32_t q_addr = smem_u32(sQ) + (uint32_t)t * q_bytes, k_addr = smem_u32(sK), v_addr = smem_u32(sV);
      const uint32_t tm_s = tmem + (uint32_t)t * kPpTileCols;
      const uint32_t tm_o = tm_s + 64u, tm_p = tm_s + 128u;
      auto issue_qk = [&](int j) {
        const int st = j % kPpStages;
        mbar_wait(&bars.k_full[st], (uint32_t)(j / kPpStages) & 1);
        tc_fence_after();
        for (int s = 0; s < k_steps; ++s) {
          const uint64_t da = make_smem_desc_sw128(q_addr + (uint32_t)s * 32u, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(k_addr + (uint32_t)st * kv_bytes + (uint32_t)s * 32u, 16, 1024);
          if (elect_one()) mma_ss(tm_s, da, db, idesc_qk, s > 0);
        }
        if (elect_one()) {
          tc_commit(&bars.k_empty[st]);                 // count 3: one arrival per tile's issuer
          tc_commit(&bars.s_full[t]);
        }
      };
      mbar_wait(&bars.q_full, 0);
      issue_qk(0);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kPpStages;
        if (j + 1 < n_tiles) {
          mbar_wait(&bars.s_free[t], (uint32_t)j & 1);    // S_t(j) is in the softmax threads' registers
          issue_qk(j + 1);
        }
        mbar_wait(&bars.v_full[st], (uint32_t)(j / kPpStages) & 1);
        mbar_wait(&bars.p_full[t], (uint32_t)j & 1);      // P_t(j) in TMEM, O_t rescaled if needed
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < BN / 16; ++s) {
          const uint64_t db = make_smem_desc_sw128(v_addr + (uint32_t)st * kv_bytes + (uint32_t)s * 2048u, kv_bytes, 1024);
          if (elect_one()) mma_ts(tm_o, tm_p + (uint32_t)s * 8u, db, idesc_pv, (j > 0) || (s > 0));
        }
        if (elect_one()) {
          tc_commit(&bars.v_empty[st]);
          tc_commit(&bars.p_empty[t]);
          if (j + 1 == n_tiles) tc_commit(&bars.o_done[t]);
        }
      }
    }
  } else {
    // =========================== softmax / correction / epilogue: twelve warps, three tiles ==========
    pp_reg_inc<152>();
    const int t = warp >> 2;                              // tile A / B / C
    const int quarter = warp & 3;                         // TMEM lane quarter = scheduler
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t tm_s = tmem + (uint32_t)t * kPpTileCols;
    const uint32_t tm_o = tm_s + 64u;
    const uint32_t tm_p = tm_s + 128u;
    const int row = (q_group * kPpTiles + t) * kPpBM + quarter * 32 + lane;
    float m_ref = 0.0f, l = 0.0f;

    for (int j = 0; j < n_tiles; ++j) {
      const bool seg2 = j >= t1;
      const int row0 = (seg2 ? j - t1 : j) * BN;
      const int valid = min(BN, (seg2 ? P.n_kv2 : P.n_kv) - row0);

      mbar_wait(&bars.s_full[t], (uint32_t)j & 1);
      tc_fence_after();
      uint32_t sr[BN];
      tmem_ld_x32(tm_s + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      tmem_ld_x32(tm_s + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      if (j > 0) {
        // P(j-1) hand-over, deferred to here so that its TMEM-store latency overlaps this tile's S load
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full[t]);
      }
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars.s_free[t]);

      if (valid < BN) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= valid) sr[i] = 0xff800000u;   // -inf
      }
      // row max (FMNMX3, four chains), rescale decision
      float mx[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) mx[u] = __uint_as_float(sr[u]);
#pragma unroll
      for (int i = 0; i < BN; i += 8)
#pragma unroll
        for (int u = 0; u < 4; ++u)
          mx[u] = fmax3(mx[u], __uint_as_float(sr[i + 2 * u]), __uint_as_float(sr[i + 2 * u + 1]));
      const float cand = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]) * P.scale_log2;
      float alpha = 1.0f;
      bool need = false;
      if (j == 0) {
        m_ref = cand;
      } else if (cand > m_ref + kPpRescaleThreshold) {
        alpha = ex2_approx(m_ref - cand);
        m_ref = cand;
        need = true;
      }

      // ---- the exponential section: this warp's turn on the XU ------------------------------------
      // Only the 64 MUFU.EX2 sit inside the turn.  The affine part (FFMA2) is done before the token arrives, the row sums
      // and the bf16 packing after it has been passed on: measured, a warp that interleaves MUFU with its dependent
      // FADD2 / F2FP (in-order issue) needs ~850 cycles for the 64 exponentials alone on its XU, against 512 for the
      // bare MUFU stream.
      const uint64_t c2 = pack2(P.scale_log2, P.scale_log2);
      float neg_m = -m_ref;
      if (P.order) {
        // turn-taking: burst n = 3 j + t of this scheduler may start once burst n - order has finished (order = how many
        // warps of a scheduler may be in their exponentials at a time).  The finished bursts of every warp are counted in
        // shared memory; the poll is a volatile load, and the reference of the affine part picks up a (zero) bit of the
        // value it read: ptxas hoists register-only work -- the whole MUFU stream -- above anything it has no data
        // dependency on, including bar.sync.
        const int tp = t - P.order >= 0 ? t - P.order : t - P.order + kPpTiles;
        const uint32_t need = (uint32_t)(t - P.order >= 0 ? j + 1 : j);
        const uint32_t addr = smem_u32(&bars.turn[tp][quarter]);
        uint32_t seen;
        do {
          asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(seen) : "r"(addr) : "memory");
        } while (seen < need);
        neg_m += __uint_as_float(seen >> 31);
      }
      const uint64_t nm2 = pack2(neg_m, neg_m);
      float ex[BN];
      // The turn is handed on EARLY, after kPpSignalAt of the 64 exponentials have been issued: the hand-over (a store, the
      // next warp's poll, its first FFMA2s) takes ~150-300 cycles, which the XU spends on the rest of this warp's stream.
      // The exponentials behind the signal take their reference through a zero loaded after the store -- the same
      // dependency trick as above, so that ptxas cannot merge them back in front of it.
#pragma unroll
      for (int i = 0; i < kPpSignalAt; i += 2) {
        const uint64_t x2 = ffma2(pack2(__uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1])), c2, nm2);
        unpack2(x2, ex[i], ex[i + 1]);
      }
#pragma unroll
      for (int i = 0; i < kPpSignalAt; ++i) ex[i] = pp_ex2(ex[i]);
      float neg_m2 = neg_m;
      if (P.order) {
        if (lane == 0) asm volatile("st.volatile.shared.u32 [%0], %1;" :: "r"(smem_u32(&bars.turn[t][quarter])), "r"(j + 1) : "memory");
        uint32_t z;
        asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(z) : "r"(smem_u32(&bars.zero)) : "memory");
        neg_m2 += __uint_as_float(z >> 31);
      }
      const uint64_t nm2b = pack2(neg_m2, neg_m2);
#pragma unroll
      for (int i = kPpSignalAt; i < BN; i += 2) {
        const uint64_t x2 = ffma2(pack2(__uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1])), c2, nm2b);
        unpack2(x2, ex[i], ex[i + 1]);
      }
#pragma unroll
      for (int i = kPpSignalAt; i < BN; ++i) ex[i] = pp_ex2(ex[i]);
      // Everything that consumes the exponentials sits in a loop that runs exactly once but whose trip count (a kernel
      // parameter) ptxas does not know: a basic-block boundary.  Inside one block ptxas threads the FADD2 / F2FP between
      // the MUFUs (no data dependency forbids it); with in-order issue every such consumer then stalls the warp until
      // its operands are back from the XU queue (experiments/mufu_rate.cu: one warp per scheduler reaches 15.7 MUFU/clk/SM
      // with a bare MUFU stream and 10.6 with the consumers threaded in).
      uint64_t acc_a = 0ull, acc_b = 0ull;
      uint32_t pk[BN / 2];
#pragma unroll
      for (int i = 0; i < BN / 2; ++i) pk[i] = 0u;
#pragma unroll 1
      for (int once = 0; once < P.one; ++once) {
#pragma unroll
        for (int i = 0; i < BN; i += 4) {
          acc_a = fadd2(acc_a, pack2(ex[i + 0], ex[i + 1]));
          acc_b = fadd2(acc_b, pack2(ex[i + 2], ex[i + 3]));
          pk[i / 2 + 0] = pack_bf16(ex[i + 0], ex[i + 1]);
          pk[i / 2 + 1] = pack_bf16(ex[i + 2], ex[i + 3]);
        }
      }

      if (j > 0) {                                         // PV(j-1) still reads P (and writes O) until its commit
        mbar_wait(&bars.p_empty[t], (uint32_t)(j - 1) & 1);
        tc_fence_after();
      }
      tmem_st_x32(tm_p + lane_off, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      float sa0, sa1, sb0, sb1;
      unpack2(acc_a, sa0, sa1);
      unpack2(acc_b, sb0, sb1);
      l = l * alpha + ((sa0 + sa1) + (sb0 + sb1));
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        for (int c = 0; c < P.d; c += 8) {
          uint32_t o8[8];
          tmem_ld_x8(tm_o + lane_off + c, o8);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
          tmem_st_x8(tm_o + lane_off + c, o8);
        }
      }
      if (j + 1 == n_tiles) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full[t]);
      }
    }

    // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------------
    mbar_wait(&bars.o_done[t], 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = P.o + ((long long)b * P.n_q + row) * P.ld_o + (long long)h * P.d;
    for (int c = 0; c < P.d; c += 8) {
      uint32_t o8[8];
      tmem_ld_x8(tm_o + lane_off + c, o8);
      tmem_wait_ld();
      if (row < P.n_q) {
        uint4 pkd;
        pkd.x = pack_bf16(__uint_as_float(o8[0]) * inv_l, __uint_as_float(o8[1]) * inv_l);
        pkd.y = pack_bf16(__uint_as_float(o8[2]) * inv_l, __uint_as_float(o8[3]) * inv_l);
        pkd.z = pack_bf16(__uint_as_float(o8[4]) * inv_l, __uint_as_float(o8[5]) * inv_l);
        pkd.w = pack_bf16(__uint_as_float(o8[6]) * inv_l, __uint_as_float(o8[7]) * inv_l);
        *reinterpret_cast<uint4*>(orow + c) = pkd;
      }
    }
    tc_fence_before();
  }

  __syncthreads();
  if (warp == kPpTmaWarp) {
    tc_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

int launch_attn_pp(const void* q, const void* k, const void* v, void* o, int batch, int heads, int n_q, int n_kv,
                   int d, long long ld_q, long long ld_k, long long ld_v, long long ld_o, float scale,
                   const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2, int order, cudaStream_t st) {
  const bool has2 = k2 != nullptr && n_kv2 > 0;
  AttnPpParams P;
  P.o = reinterpret_cast<__nv_bfloat16*>(o);
  P.ld_o = ld_o;
  P.heads = heads; P.n_q = n_q; P.n_kv = n_kv; P.n_kv2 = has2 ? n_kv2 : 0;
  P.d = d; P.d_pad = (d + 15) / 16 * 16;
  P.scale_log2 = scale * 1.4426950408889634f;
  P.order = order;
  P.one = 1;
  if (P.d_pad > 64) return fail("vf_attn_fwd(bf16, round-robin arrangement): d_head=%d > 64", d);
  CUtensorMap mq, mk, mv, mk2, mv2;
  if (int rc = attn_make_map(&mq, q, batch, heads, n_q, d, ld_q, kPpBM)) return rc;
  if (int rc = attn_make_map(&mk, k, batch, heads, n_kv, d, ld_k, kPpBN)) return rc;
  if (int rc = attn_make_map(&mv, v, batch, heads, n_kv, d, ld_v, kPpBN)) return rc;
  if (has2) {
    if (int rc = attn_make_map(&mk2, k2, batch, heads, n_kv2, d, ld_k2, kPpBN)) return rc;
    if (int rc = attn_make_map(&mv2, v2, batch, heads, n_kv2, d, ld_v2, kPpBN)) return rc;
  } else {
    mk2 = mk;
    mv2 = mv;
  }
  const size_t smem = 1024 + (size_t)kPpTiles * kPpBM * 128 + (size_t)2 * kPpStages * kPpBN * 128;
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(attn_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  dim3 grid((n_q + kPpTiles * kPpBM - 1) / (kPpTiles * kPpBM), batch * heads);
  attn_pp_kernel<<<grid, kPpThreads, smem, st>>>(mq, mk, mv, mk2, mv2, P);
  return check_cuda(cudaGetLastError(), "attn_pp_kernel launch");
}

}  // namespace vf
