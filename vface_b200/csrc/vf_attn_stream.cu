// Fused attention core, "streamed softmax" arrangement (dtype VF_BF16 of vf_attn_fwd, d_head <= 128).
//
//   o = softmax(q k^T * scale) v   per (batch, head), q/k/v/o in the reference's native (batch, n, heads*d) layout
//   (ldm/models/pnp_utils.py:270-286, ldm/modules/attention.py:203-220), optional second K/V segment.
//
// Why a second kernel.  At d_head = 40 the kernel owes one exponential per 160 MMA flops, so the MUFU (XU) pipe, not the
// tensor pipe, is the binding unit.  The first arrangement (vf_attn_tc.cu) left the XU idle ~25 % of the
// time: a softmax warp stopped issuing MUFU at every tile boundary (barrier round trip, TMEM load, P hand-over, all
// queued behind the other warps' MUFU traffic in the MIO path), and the three warps per scheduler only covered each
// other's gaps by chance.  Here every softmax warp keeps its own MUFU stream running ACROSS tile boundaries:
//
//   * S is double-buffered in TMEM (S0 | S1 | O = 64 + 64 + d_pad columns of one 256-column allocation, two CTAs per
//     SM); the issuer runs QK two tiles ahead: PV_j then QK_{j+2}, so S_{j+1} is complete long before the softmax
//     warps finish tile j.
//   * a warp walks a tile in two 32-column halves and always has the NEXT half in flight: the second half of tile j
//     is loaded (tcgen05.ld) while the first is in its exponentials, the first half of tile j+1 while the second is.
//     The s_full wait of tile j+1 is therefore issued in the middle of tile j, on a barrier that completed long ago.
//   * P (bf16) is written over the consumed first half of the same S buffer (one tcgen05.st per tile), so there is no
//     separate P allocation and no P-empty barrier: PV_j reads P_j from buffer j & 1, QK_{j+2} overwrites that buffer
//     behind it in the tensor pipe's issue order.  Two barrier operations per tile per softmax warp (s_full wait,
//     p_full arrive; the arrive is deferred behind the next tile's first TMEM load so the tcgen05.st latency overlaps).
//   * the row max is off the critical path (lagged reference): the exponentials of tile j use the reference decided
//     from tiles < j, the max of tile j is computed in the shadow of its MUFU stream (FMNMX3 on the ALU pipe) and only
//     decides the reference of tile j+1; O and l are rescaled lazily (reference moved by > 2^8).  A row whose scores
//     jump by more than 2^64 between tiles takes the exact path for that tile (S_j is still intact in TMEM because P_j
//     is stored only after the decision).  The result is the same softmax.
//
// Warps: 0-3 softmax (one thread per query row = TMEM lane), 4 TMA producer (K runs two tiles ahead of V), 7 MMA issuer.
#include "vf_attn.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cstdlib>

namespace vf {

using namespace sm100;

namespace {

constexpr int kThreads = 256;
constexpr int kTmaWarp = 4;
constexpr int kMmaWarp = 7;
constexpr int kBM = 128;                     // query rows per CTA
constexpr int kBN = 64;                      // keys per tile
constexpr int kTmemCols = 256;               // S0 [0,64) | S1 [64,128) | O [128, 128 + d_pad)
constexpr int kMaxStages = 3;
constexpr float kRescaleThreshold = 8.0f;    // log2 units the reference may lag before O/l are rescaled
constexpr float kLagGuard = 64.0f;           // log2 units a row may exceed its lagged reference before the exact path

struct Params {
  __nv_bfloat16* o;
  long long ld_o;
  int heads, n_q, n_kv, n_kv2, d, d_pad, kb;   // kb = ceil(d / 64) 64-wide head-dim blocks
  float scale_log2;                            // scale * log2(e)
};

struct __align__(8) Barriers {
  uint64_t q_full;
  uint64_t k_full[kMaxStages], k_empty[kMaxStages];
  uint64_t v_full[kMaxStages], v_empty[kMaxStages];
  uint64_t s_full[2];      // QK_j complete -> S buffer j & 1 (tcgen05.commit, count 1)
  uint64_t p_full[2];      // P_j stored over S buffer j & 1 and O rescaled if needed (count 4: one arrive per softmax warp)
  uint64_t pv_done;        // PV_j complete (tcgen05.commit); only the rare O-rescale path waits on it
  uint64_t o_done;
  uint32_t tmem_base;
};

template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kRegs)); }

// tcgen05.wait::ld that also names the destination registers of the loads it waits for, so that no use of them can be
// scheduled in front of it.
__device__ __forceinline__ void wait_ld_32(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                 "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                 "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :: "memory");
}

// kEmu of every 4 score pairs take their exp2 on the FMA pipe (Cody-Waite + degree-3 polynomial) instead of MUFU.
template <int kStages, int kEmu>
__global__ void __launch_bounds__(kThreads, 2)
attn_stream_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                   const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_k2,
                   const __grid_constant__ CUtensorMap map_v2, const Params P) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ Barriers bars;

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_tile = blockIdx.x;
  const int bh = blockIdx.y;
  const int b = bh / P.heads, h = bh - b * P.heads;

  // ---- shared memory carve-up (1024-byte aligned tiles for the 128B swizzle) -------------------
  const uint32_t dyn_base = smem_u32(smem_dyn);
  const uint32_t tile_base = (dyn_base + 1023u) & ~1023u;
  unsigned char* tiles = smem_dyn + (tile_base - dyn_base);
  const uint32_t q_block_bytes = kBM * 128;
  const uint32_t kv_block_bytes = kBN * 128;
  const uint32_t q_bytes = P.kb * q_block_bytes;
  const uint32_t kv_bytes = P.kb * kv_block_bytes;
  unsigned char* sQ = tiles;
  unsigned char* sK = sQ + q_bytes;                 // kStages stages
  unsigned char* sV = sK + kStages * kv_bytes;      // kStages stages

  const int t1 = (P.n_kv + kBN - 1) / kBN;
  const int t2 = (P.n_kv2 + kBN - 1) / kBN;
  const int n_tiles = t1 + t2;

  if (threadIdx.x == 0) {
    mbar_init(&bars.q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars.k_full[s], 1);
      mbar_init(&bars.k_empty[s], 1);
      mbar_init(&bars.v_full[s], 1);
      mbar_init(&bars.v_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars.s_full[s], 1);
      mbar_init(&bars.p_full[s], 4);
    }
    mbar_init(&bars.pv_done, 1);
    mbar_init(&bars.o_done, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) tmem_alloc<kTmemCols>(&bars.tmem_base);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  const uint32_t tm_o = tmem + 2 * kBN;

  if (warp >= 4) {
    reg_dec<40>();
    if (warp == kTmaWarp) {
      // =========================== TMA producer ==================================================
      // K runs two tiles ahead of V (the issuer runs QK two tiles ahead of PV): K_0, K_1, then V_j, K_{j+2}.
      if (lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_arrive_expect_tx(&bars.q_full, q_bytes);
        for (int kb = 0; kb < P.kb; ++kb)
          tma_load_4d(sQ + kb * q_block_bytes, &map_q, &bars.q_full, kb * 64, h, q_tile * kBM, b);
        auto load = [&](int j, bool is_v) {
          const int st = j % kStages;
          const uint32_t use = (uint32_t)(j / kStages);
          const bool seg2 = j >= t1;
          const int row0 = (seg2 ? j - t1 : j) * kBN;
          const CUtensorMap* m = is_v ? (seg2 ? &map_v2 : &map_v) : (seg2 ? &map_k2 : &map_k);
          uint64_t* full = is_v ? &bars.v_full[st] : &bars.k_full[st];
          uint64_t* empty = is_v ? &bars.v_empty[st] : &bars.k_empty[st];
          unsigned char* dst = (is_v ? sV : sK) + st * kv_bytes;
          mbar_wait(empty, (use & 1) ^ 1);
          mbar_arrive_expect_tx(full, kv_bytes);
          for (int kb = 0; kb < P.kb; ++kb)
            tma_load_4d(dst + kb * kv_block_bytes, m, full, kb * 64, h, row0, b);
        };
        load(0, false);
        if (n_tiles > 1) load(1, false);
        for (int j = 0; j < n_tiles; ++j) {
          load(j, true);
          if (j + 2 < n_tiles) load(j + 2, false);
        }
      }
    } else if (warp == kMmaWarp) {
      // =========================== MMA issuer ====================================================
      // Warp-uniform loop (all lanes wait on the barriers, one elected lane issues): descriptors stay on the uniform
      // datapath.  Issue order: QK_0, QK_1, then per tile j: PV_j (P_j in S buffer j & 1), QK_{j+2} into that buffer.
      const uint32_t idesc_qk = make_idesc_bf16(kBM, kBN, false);
      const uint32_t idesc_pv = make_idesc_bf16(kBM, P.d_pad, true);
      const int k_steps = P.d_pad / 16;
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t k_addr = smem_u32(sK);
      const uint32_t v_addr = smem_u32(sV);

      auto issue_qk = [&](int j) {
        const int st = j % kStages;
        const uint32_t tm_s = tmem + (uint32_t)(j & 1) * kBN;
        mbar_wait(&bars.k_full[st], (uint32_t)(j / kStages) & 1);
        tc_fence_after();
        for (int s = 0; s < k_steps; ++s) {
          const uint32_t off_blk = (uint32_t)(s >> 2), off_in = (uint32_t)(s & 3) * 32u;
          const uint64_t da = make_smem_desc_sw128(q_addr + off_blk * q_block_bytes + off_in, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(k_addr + st * kv_bytes + off_blk * kv_block_bytes + off_in, 16, 1024);
          if (elect_one()) mma_ss(tm_s, da, db, idesc_qk, s > 0);
        }
        if (elect_one()) {
          tc_commit(&bars.k_empty[st]);
          tc_commit(&bars.s_full[j & 1]);
        }
      };

      mbar_wait(&bars.q_full, 0);
      issue_qk(0);
      if (n_tiles > 1) issue_qk(1);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStages;
        const uint32_t tm_p = tmem + (uint32_t)(j & 1) * kBN;            // P_j over the first 32 columns of S_j
        mbar_wait(&bars.v_full[st], (uint32_t)(j / kStages) & 1);
        mbar_wait(&bars.p_full[j & 1], (uint32_t)(j >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < kBN / 16; ++s) {
          // B = V tile, MN-major: 64 head-dim elements contiguous (128 B) per key row, 8-row groups 1024 B apart (SBO),
          // further 64-wide head-dim blocks kv_block_bytes apart (LBO).
          const uint64_t db = make_smem_desc_sw128(v_addr + st * kv_bytes + (uint32_t)s * 2048u, kv_block_bytes, 1024);
          if (elect_one()) mma_ts(tm_o, tm_p + (uint32_t)s * 8u, db, idesc_pv, (j > 0) || (s > 0));
        }
        if (elect_one()) {
          tc_commit(&bars.v_empty[st]);
          tc_commit(&bars.pv_done);
          if (j + 1 == n_tiles) tc_commit(&bars.o_done);
        }
        if (j + 2 < n_tiles) issue_qk(j + 2);        // behind PV_j in the pipe: may overwrite P_j
      }
    }
  } else {
    // =========================== softmax / correction / epilogue ================================
    reg_inc<216>();
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;       // TMEM lane quarter this warp may touch
    const int row = q_tile * kBM + warp * 32 + lane;             // query row owned by this thread
    const uint64_t c2 = pack2(P.scale_log2, P.scale_log2);

    float m_ref = 0.0f, l = 0.0f;
    float lag_alpha = 1.0f, lag_m = 0.0f;     // rescale decided by the previous tile's row max
    bool lag_need = false;

    uint32_t cur[32], nxt[32];                // first / second 32-column half of the tile being processed
    uint32_t pk[32];                          // P_j as bf16x2
    float mx[4];
    uint64_t acc_a, acc_b;

    auto mask_half = [&](uint32_t (&s)[32], int valid) {          // columns >= valid are out of range
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (i >= valid) s[i] = 0xff800000u;   // -inf
    };
    auto max_half = [&](const uint32_t (&s)[32]) {
#pragma unroll
      for (int i = 0; i < 32; i += 8)
#pragma unroll
        for (int t = 0; t < 4; ++t)
          mx[t] = fmax3(mx[t], __uint_as_float(s[i + 2 * t]), __uint_as_float(s[i + 2 * t + 1]));
    };
    // p = exp2(s * c - m_ref) for one half: packed FFMA2 for the affine part, MUFU.EX2 (or the FMA-pipe polynomial) per
    // element, packed FADD2 row sums in two chains, bf16x2 packing, and -- kMax -- this half's row max (FMNMX3, four chains).
    auto exp_half = [&](const uint32_t (&s)[32], uint32_t* out, const bool with_max) {
      const uint64_t nm2 = pack2(-m_ref, -m_ref);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const uint64_t xa = ffma2(pack2(__uint_as_float(s[i + 0]), __uint_as_float(s[i + 1])), c2, nm2);
        const uint64_t xb = ffma2(pack2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), c2, nm2);
        if (with_max) {
          mx[(i / 4) & 3] = fmax3(mx[(i / 4) & 3], __uint_as_float(s[i + 0]), __uint_as_float(s[i + 1]));
          mx[(i / 4 + 2) & 3] = fmax3(mx[(i / 4 + 2) & 3], __uint_as_float(s[i + 2]), __uint_as_float(s[i + 3]));
        }
        float p0, p1, p2, p3;
        if (((i / 2) & 3) < kEmu) {
          exp2_poly2(xa, p0, p1);
        } else {
          float t0, t1;
          unpack2(xa, t0, t1);
          p0 = ex2_approx(t0); p1 = ex2_approx(t1);
        }
        if (((i / 2 + 1) & 3) < kEmu) {
          exp2_poly2(xb, p2, p3);
        } else {
          float t2, t3;
          unpack2(xb, t2, t3);
          p2 = ex2_approx(t2); p3 = ex2_approx(t3);
        }
        acc_a = fadd2(acc_a, pack2(p0, p1));
        acc_b = fadd2(acc_b, pack2(p2, p3));
        out[i / 2 + 0] = pack_bf16(p0, p1);
        out[i / 2 + 1] = pack_bf16(p2, p3);
      }
    };
    auto reset_max = [&]() {
#pragma unroll
      for (int t = 0; t < 4; ++t) mx[t] = __uint_as_float(0xff800000u);
    };
    auto row_cand = [&]() { return fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]) * P.scale_log2; };

    // ---- prologue: S_0, both halves ------------------------------------------------------------------
    mbar_wait(&bars.s_full[0], 0);
    tc_fence_after();
    tmem_ld_x32(tmem + lane_off, cur);

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t tm_s = tmem + (uint32_t)(j & 1) * kBN;
      const bool seg2 = j >= t1;
      const int row0 = (seg2 ? j - t1 : j) * kBN;
      const int valid = min(kBN, (seg2 ? P.n_kv2 : P.n_kv) - row0);

      // tcgen05.wait::ld covers EVERY load issued so far: retire the first half (issued in the middle of the previous tile)
      // before the second half goes in flight, so that the latter overlaps the first half's exponentials
      if (j > 0) wait_ld_32(cur);
      tmem_ld_x32(tm_s + lane_off + 32, nxt);
      if (j > 0) {
        // P_{j-1} hand-over, deferred to here so that its tcgen05.st latency overlaps the load just issued
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full[(j - 1) & 1]);
      }
      float alpha = lag_alpha;              // (j == 0: 1, false)
      bool need = lag_need;
      acc_a = 0ull; acc_b = 0ull;
      reset_max();
      bool redo = false;
      if (j == 0) {
        // no reference yet: exact order for this tile -- max over both halves first
        wait_ld_32(cur);
        wait_ld_32(nxt);
        if (valid < kBN) { mask_half(cur, valid); mask_half(nxt, valid - 32); }
        max_half(cur);
        max_half(nxt);
        m_ref = row_cand();
        exp_half(cur, &pk[0], false);
        if (n_tiles > 1) {
          mbar_wait(&bars.s_full[1], 0);
          tc_fence_after();
          tmem_ld_x32(tmem + kBN + lane_off, cur);          // first half of S_1
        }
        exp_half(nxt, &pk[16], false);
      } else {
        if (valid < kBN) mask_half(cur, valid);
        exp_half(cur, &pk[0], true);
        wait_ld_32(nxt);
        if (valid < kBN) mask_half(nxt, valid - 32);
        if (j + 1 < n_tiles) {
          // first half of S_{j+1}: QK_{j+1} was issued behind PV_{j-1}, a full tile ago
          mbar_wait(&bars.s_full[(j + 1) & 1], (uint32_t)((j + 1) >> 1) & 1);
          tc_fence_after();
          tmem_ld_x32(tmem + (uint32_t)((j + 1) & 1) * kBN + lane_off, cur);
        }
        exp_half(nxt, &pk[16], true);
        const float cand = row_cand();
        redo = __any_sync(0xffffffffu, cand > m_ref + kLagGuard);
        if (redo) {
          // exact path for this tile: S_j is still intact in TMEM (P_j has not been stored yet).  `cur` is busy with the
          // load of S_{j+1}, so both halves go through `nxt`.
          if (cand > m_ref + kRescaleThreshold) {
            alpha *= ex2_approx(m_ref - cand);
            m_ref = cand;
            need = true;
          }
          acc_a = 0ull; acc_b = 0ull;
          if (j + 1 < n_tiles) wait_ld_32(cur);
          tmem_ld_x32(tm_s + lane_off, nxt);
          wait_ld_32(nxt);
          if (valid < kBN) mask_half(nxt, valid);
          exp_half(nxt, &pk[0], false);
          tmem_ld_x32(tm_s + lane_off + 32, nxt);
          wait_ld_32(nxt);
          if (valid < kBN) mask_half(nxt, valid - 32);
          exp_half(nxt, &pk[16], false);
        }
        lag_alpha = 1.0f;
        lag_need = false;
        if (cand > m_ref + kRescaleThreshold) {       // decided now, applied to tile j+1 (its exps, then O and l)
          lag_alpha = ex2_approx(m_ref - cand);
          lag_m = cand;
          lag_need = true;
        }
      }

      // P_j over the consumed first half of S_j
      tmem_st_x32(tm_s + lane_off, pk);
      float sa0, sa1, sb0, sb1;
      unpack2(acc_a, sa0, sa1);
      unpack2(acc_b, sb0, sb1);
      l = l * alpha + ((sa0 + sa1) + (sb0 + sb1));
      if (lag_need) m_ref = lag_m;                    // reference of the next tile
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rare: the reference moved.  O holds PV_0..PV_{j-1} in the old reference; PV_{j-1} must have retired.
        mbar_wait(&bars.pv_done, (uint32_t)(j - 1) & 1);
        tc_fence_after();
        for (int c = 0; c < P.d_pad; c += 8) {
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
        if (lane == 0) mbar_arrive(&bars.p_full[j & 1]);
      }
    }

    // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------------
    mbar_wait(&bars.o_done, 0);
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
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem);
  }
}

template <int kStages, int kEmu>
int launch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mk2,
           const CUtensorMap& mv2, const Params& P, int batch, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)P.kb * (kBM * 128 + 2 * kStages * kBN * 128);
  static bool attr = false;        // one process drives one GPU (vf_capi.cu: check_device binds the library to it)
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(attn_stream_kernel<kStages, kEmu>, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
    attr = true;
  }
  dim3 grid((P.n_q + kBM - 1) / kBM, batch * P.heads);
  attn_stream_kernel<kStages, kEmu><<<grid, kThreads, smem, st>>>(mq, mk, mv, mk2, mv2, P);
  return check_cuda(cudaGetLastError(), "attn_stream_kernel launch");
}

}  // namespace

int attn_make_map(CUtensorMap* m, const void* base, int batch, int heads, int n, int d, long long ld, int box_rows);

// d_pad <= 128.  Pointers, strides and shapes have been validated by launch_attn_tc.
int launch_attn_stream(const void* q, const void* k, const void* v, void* o, int batch, int heads, int n_q, int n_kv,
                       int d, long long ld_q, long long ld_k, long long ld_v, long long ld_o, float scale,
                       const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2, int emu,
                       cudaStream_t st) {
  const bool has2 = k2 != nullptr && n_kv2 > 0;
  Params P;
  P.o = reinterpret_cast<__nv_bfloat16*>(o);
  P.ld_o = ld_o;
  P.heads = heads; P.n_q = n_q; P.n_kv = n_kv; P.n_kv2 = has2 ? n_kv2 : 0;
  P.d = d; P.d_pad = (d + 15) / 16 * 16; P.kb = (d + 63) / 64;
  P.scale_log2 = scale * 1.4426950408889634f;
  CUtensorMap mq, mk, mv, mk2, mv2;
  if (int rc = attn_make_map(&mq, q, batch, heads, n_q, d, ld_q, kBM)) return rc;
  if (int rc = attn_make_map(&mk, k, batch, heads, n_kv, d, ld_k, kBN)) return rc;
  if (int rc = attn_make_map(&mv, v, batch, heads, n_kv, d, ld_v, kBN)) return rc;
  if (has2) {
    if (int rc = attn_make_map(&mk2, k2, batch, heads, n_kv2, d, ld_k2, kBN)) return rc;
    if (int rc = attn_make_map(&mv2, v2, batch, heads, n_kv2, d, ld_v2, kBN)) return rc;
  } else {
    mk2 = mk;
    mv2 = mv;
  }
  if (P.kb == 1) {                       // 16 KB Q + 3 x 16 KB K/V: 65 KB per CTA
    switch (emu) {
      case 1: return launch<3, 1>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 2: return launch<3, 2>(mq, mk, mv, mk2, mv2, P, batch, st);
      default: return launch<3, 0>(mq, mk, mv, mk2, mv2, P, batch, st);
    }
  }
  return launch<2, 0>(mq, mk, mv, mk2, mv2, P, batch, st);      // kb = 2: 32 KB Q + 2 x 32 KB K/V: 97 KB per CTA
}

}  // namespace vf
