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
constexpr int kTmemCols = 256;               // kSBuf S buffers of 64 columns | O [kSBuf * 64, kSBuf * 64 + d_pad)
constexpr int kMaxStages = 3;
constexpr float kRescaleThreshold = 8.0f;    // log2 units the reference may lag before O/l are rescaled
constexpr float kLagGuard = 64.0f;           // log2 units a row may exceed its lagged reference before the exact path

struct Params {
  __nv_bfloat16* o;
  long long ld_o;
  int heads, n_q, n_kv, n_kv2, d, d_pad, kb;   // kb = ceil(d / 64) 64-wide head-dim blocks of Q and K (the QK^T reduction)
  // Wide heads (the first-stage AttnBlock: ONE head of width 512, ldm/modules/diffusionmodules/model.py:150-203): the
  // scores reduce over all d columns of Q/K, but one CTA accumulates only a dv-wide column slice of O (TMEM holds
  // 2 x 64 S columns + dv_pad O columns).  `heads` counts those slices (CTA index), vgroup of them share one Q/K head.
  int dv, dv_pad, kbv, vgroup;
  float scale_log2;                            // scale * log2(e)
};

struct __align__(8) Barriers {
  uint64_t q_full;
  uint64_t k_full[kMaxStages], k_empty[kMaxStages];
  uint64_t v_full[kMaxStages], v_empty[kMaxStages];
  uint64_t s_full[3];      // QK_j complete -> S buffer j % kSBuf (tcgen05.commit, count 1)
  uint64_t p_full[3];      // P_j stored over S buffer j % kSBuf and O rescaled if needed (count 4: one arrive per softmax warp)
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
// kSBuf S buffers in TMEM: the issuer runs QK kSBuf tiles ahead of PV (PV_j, then QK_{j+kSBuf} into the buffer P_j sat in).
//   With two, S_{j+1} depends on P_{j-1}, which the softmax warps hand over at the top of tile j, and is wanted in the middle
//   of tile j: half a tile of slack against an issue + MMA + two barrier round trips of ~600 cycles -- measured: every
//   softmax warp stalls there and the XU idles half the time (369 TF/s).  With three (d_pad <= 64: 3 x 64 + 48 = 240
//   columns) the slack is a tile and a half.
// kLateArrive: the p_full hand-over of tile j-1 waits for its tcgen05.st in the MIDDLE of tile j instead of at its top.
template <int kStages, int kEmu, int kSBuf, bool kLateArrive>
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
  const uint32_t k_bytes = P.kb * kv_block_bytes;
  const uint32_t v_bytes = P.kbv * kv_block_bytes;
  const int hq = h / P.vgroup;                      // Q/K head of this O column slice
  unsigned char* sQ = tiles;
  unsigned char* sK = sQ + q_bytes;                 // kStages stages
  unsigned char* sV = sK + kStages * k_bytes;       // kStages stages

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
    for (int s = 0; s < kSBuf; ++s) {
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
  const uint32_t tm_o = tmem + kSBuf * kBN;

  if (warp >= 4) {
    reg_dec<40>();
    if (warp == kTmaWarp) {
      // =========================== TMA producer ==================================================
      // K runs kSBuf tiles ahead of V (the issuer runs QK that far ahead of PV): K_0 .. K_{kSBuf-1}, then V_j, K_{j+kSBuf}.
      if (lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_arrive_expect_tx(&bars.q_full, q_bytes);
        for (int kb = 0; kb < P.kb; ++kb)
          tma_load_4d(sQ + kb * q_block_bytes, &map_q, &bars.q_full, kb * 64, hq, q_tile * kBM, b);
        auto load = [&](int j, bool is_v) {
          const int st = j % kStages;
          const uint32_t use = (uint32_t)(j / kStages);
          const bool seg2 = j >= t1;
          const int row0 = (seg2 ? j - t1 : j) * kBN;
          const CUtensorMap* m = is_v ? (seg2 ? &map_v2 : &map_v) : (seg2 ? &map_k2 : &map_k);
          uint64_t* full = is_v ? &bars.v_full[st] : &bars.k_full[st];
          uint64_t* empty = is_v ? &bars.v_empty[st] : &bars.k_empty[st];
          unsigned char* dst = is_v ? sV + st * v_bytes : sK + st * k_bytes;
          mbar_wait(empty, (use & 1) ^ 1);
          mbar_arrive_expect_tx(full, is_v ? v_bytes : k_bytes);
          const int nkb = is_v ? P.kbv : P.kb;
          for (int kb = 0; kb < nkb; ++kb)
            tma_load_4d(dst + kb * kv_block_bytes, m, full, kb * 64, is_v ? h : hq, row0, b);
        };
        for (int j = 0; j < kSBuf && j < n_tiles; ++j) load(j, false);
        for (int j = 0; j < n_tiles; ++j) {
          load(j, true);
          if (j + kSBuf < n_tiles) load(j + kSBuf, false);
        }
      }
    } else if (warp == kMmaWarp) {
      // =========================== MMA issuer ====================================================
      // Warp-uniform loop (all lanes wait on the barriers, one elected lane issues): descriptors stay on the uniform
      // datapath.  Issue order: QK_0 .. QK_{kSBuf-1}, then per tile j: PV_j (P_j in S buffer j % kSBuf), QK_{j+kSBuf} into it.
      const uint32_t idesc_qk = make_idesc_bf16(kBM, kBN, false);
      const uint32_t idesc_pv = make_idesc_bf16(kBM, P.dv_pad, true);
      const int k_steps = P.d_pad / 16;
      const uint32_t q_addr = smem_u32(sQ);
      const uint32_t k_addr = smem_u32(sK);
      const uint32_t v_addr = smem_u32(sV);

      auto issue_qk = [&](int j) {
        const int st = j % kStages;
        const uint32_t tm_s = tmem + (uint32_t)(j % kSBuf) * kBN;
        mbar_wait(&bars.k_full[st], (uint32_t)(j / kStages) & 1);
        tc_fence_after();
        for (int s = 0; s < k_steps; ++s) {
          const uint32_t off_blk = (uint32_t)(s >> 2), off_in = (uint32_t)(s & 3) * 32u;
          const uint64_t da = make_smem_desc_sw128(q_addr + off_blk * q_block_bytes + off_in, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(k_addr + st * k_bytes + off_blk * kv_block_bytes + off_in, 16, 1024);
          if (elect_one()) mma_ss(tm_s, da, db, idesc_qk, s > 0);
        }
        if (elect_one()) {
          tc_commit(&bars.k_empty[st]);
          tc_commit(&bars.s_full[j % kSBuf]);
        }
      };

      mbar_wait(&bars.q_full, 0);
      for (int j = 0; j < kSBuf && j < n_tiles; ++j) issue_qk(j);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j % kStages;
        const uint32_t tm_p = tmem + (uint32_t)(j % kSBuf) * kBN;        // P_j over the first 32 columns of S_j
        mbar_wait(&bars.v_full[st], (uint32_t)(j / kStages) & 1);
        mbar_wait(&bars.p_full[j % kSBuf], (uint32_t)(j / kSBuf) & 1);
        tc_fence_after();
#pragma unroll
        for (int s = 0; s < kBN / 16; ++s) {
          // B = V tile, MN-major: 64 head-dim elements contiguous (128 B) per key row, 8-row groups 1024 B apart (SBO),
          // further 64-wide head-dim blocks kv_block_bytes apart (LBO).
          const uint64_t db = make_smem_desc_sw128(v_addr + st * v_bytes + (uint32_t)s * 2048u, kv_block_bytes, 1024);
          if (elect_one()) mma_ts(tm_o, tm_p + (uint32_t)s * 8u, db, idesc_pv, (j > 0) || (s > 0));
        }
        if (elect_one()) {
          tc_commit(&bars.v_empty[st]);
          tc_commit(&bars.pv_done);
          if (j + 1 == n_tiles) tc_commit(&bars.o_done);
        }
        if (j + kSBuf < n_tiles) issue_qk(j + kSBuf);        // behind PV_j in the pipe: may overwrite P_j
      }
    }
  } else {
    // =========================== softmax / correction / epilogue ================================
    reg_inc<216>();
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;       // TMEM lane quarter this warp may touch
    const int row = q_tile * kBM + warp * 32 + lane;             // query row owned by this thread
    const uint64_t c2 = pack2(P.scale_log2, P.scale_log2);

    float m_ref = 0.0f, l = 0.0f;
    float lag_alpha = 1.0f, lag_m = 0.0f;     // rescale decided at the end of the previous tile
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
    // element, packed FADD2 row sums in two chains, bf16x2 packing.
    auto exp_half = [&](const uint32_t (&s)[32], uint32_t* out) {
      const uint64_t nm2 = pack2(-m_ref, -m_ref);
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        const uint64_t xa = ffma2(pack2(__uint_as_float(s[i + 0]), __uint_as_float(s[i + 1])), c2, nm2);
        const uint64_t xb = ffma2(pack2(__uint_as_float(s[i + 2]), __uint_as_float(s[i + 3])), c2, nm2);
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
    auto tile_sum = [&]() {
      float sa0, sa1, sb0, sb1;
      unpack2(acc_a, sa0, sa1);
      unpack2(acc_b, sb0, sb1);
      return (sa0 + sa1) + (sb0 + sb1);
    };
    // exact order for one tile (no usable reference): row max over both halves, then the exponentials.  Both halves go
    // through `nxt` (`cur` may be busy with the load of the next tile); S_j is intact in TMEM (P_j not stored yet).
    auto exact_tile = [&](uint32_t tm_s, int valid, float& alpha, bool& need, bool first) {
#pragma unroll
      for (int t = 0; t < 4; ++t) mx[t] = __uint_as_float(0xff800000u);
      for (int hf = 0; hf < 2; ++hf) {
        tmem_ld_x32(tm_s + lane_off + 32 * hf, nxt);
        wait_ld_32(nxt);
        if (valid < kBN) mask_half(nxt, valid - 32 * hf);
        max_half(nxt);
      }
      const float cand = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]) * P.scale_log2;
      if (first) {
        m_ref = cand;
      } else if (cand > m_ref + kRescaleThreshold) {
        alpha *= ex2_approx(m_ref - cand);
        m_ref = cand;
        need = true;
      }
      acc_a = 0ull; acc_b = 0ull;
      for (int hf = 0; hf < 2; ++hf) {
        tmem_ld_x32(tm_s + lane_off + 32 * hf, nxt);
        wait_ld_32(nxt);
        if (valid < kBN) mask_half(nxt, valid - 32 * hf);
        exp_half(nxt, &pk[16 * hf]);
      }
    };

    // buffer / barrier-phase bookkeeping without divisions: tile j sits in S buffer sb (= j % kSBuf), phase sph (= j / kSBuf & 1)
    int sb = 0, pb = 0;
    uint32_t sph = 0;

    // ---- prologue: first half of S_0 in flight -------------------------------------------------------
    mbar_wait(&bars.s_full[0], 0);
    tc_fence_after();
    tmem_ld_x32(tmem + lane_off, cur);

    for (int j = 0; j < n_tiles; ++j) {
      const uint32_t tm_s = tmem + (uint32_t)sb * kBN;
      const bool seg2 = j >= t1;
      const int row0 = (seg2 ? j - t1 : j) * kBN;
      const int valid = min(kBN, (seg2 ? P.n_kv2 : P.n_kv) - row0);
      int nb = sb + 1;                               // S buffer and phase of tile j + 1
      uint32_t nph = sph;
      if (nb == kSBuf) { nb = 0; nph ^= 1u; }
      const bool has_next = j + 1 < n_tiles;

      auto hand_over_prev = [&]() {      // P_{j-1} (and a rescaled O) are in TMEM: PV_{j-1} may go
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full[pb]);
      };
      auto prefetch_next = [&]() {       // first half of S_{j+1}: QK_{j+1} was issued behind PV_{j+1-kSBuf}
        mbar_wait(&bars.s_full[nb], nph);
        tc_fence_after();
        tmem_ld_x32(tmem + (uint32_t)nb * kBN + lane_off, cur);
      };

      float alpha = lag_alpha;              // (j == 0: 1, false)
      bool need = lag_need;
      if (j == 0) {
        // no reference yet: exact order for this tile
        wait_ld_32(cur);                    // (retire the prologue's load; the exact path reloads both halves)
        exact_tile(tm_s, valid, alpha, need, true);
        if (has_next) prefetch_next();
      } else {
        // tcgen05.wait::ld covers EVERY load issued so far: retire the first half (issued in the middle of the previous
        // tile) before the second half goes in flight, so that the latter overlaps the first half's exponentials
        wait_ld_32(cur);
        tmem_ld_x32(tm_s + lane_off + 32, nxt);
        // P_{j-1} hand-over, deferred to here so that its tcgen05.st latency overlaps the load just issued (kLateArrive: the
        // first half's exponentials as well)
        if (!(kLateArrive && j > 1)) hand_over_prev();
        acc_a = 0ull; acc_b = 0ull;
        if (valid < kBN) mask_half(cur, valid);
        exp_half(cur, &pk[0]);
        if (kLateArrive && j > 1) hand_over_prev();
        wait_ld_32(nxt);
        if (valid < kBN) mask_half(nxt, valid - 32);
        if (has_next) prefetch_next();
        exp_half(nxt, &pk[16]);
        // The row max never enters the steady state: with every p = 2^(x - m_ref) summed anyway, the tile sum IS the
        // overflow detector (a score more than 2^kLagGuard above the reference makes it exceed 2^kLagGuard, inf or NaN),
        // and m_ref + log2(sum) >= the tile's true max is as good a new reference as the max itself.
        if (__any_sync(0xffffffffu, !(tile_sum() <= 18446744073709551616.0f))) {     // 2^64; also catches inf / NaN
          if (has_next) wait_ld_32(cur);
          exact_tile(tm_s, valid, alpha, need, false);
        }
      }
      const float ts = tile_sum();
      lag_alpha = 1.0f;
      lag_need = false;
      if (ts > 256.0f) {                         // the reference lags by more than 2^8: re-reference from the next tile on
        lag_m = m_ref + lg2_approx(ts);
        lag_alpha = ex2_approx(m_ref - lag_m);
        lag_need = true;
      }

      // P_j over the consumed first half of S_j
      tmem_st_x32(tm_s + lane_off, pk);
      l = l * alpha + ts;
      if (lag_need) m_ref = lag_m;                    // reference of the next tile
      if (j > 0 && __any_sync(0xffffffffu, need)) {
        // rare: the reference moved.  O holds PV_0..PV_{j-1} in the old reference; PV_{j-1} must have retired.
        if (kLateArrive && j == 1) { /* handed over at the top */ }
        mbar_wait(&bars.pv_done, (uint32_t)(j - 1) & 1);
        tc_fence_after();
        for (int c = 0; c < P.dv_pad; c += 8) {
          uint32_t o8[8];
          tmem_ld_x8(tm_o + lane_off + c, o8);
          tmem_wait_ld();
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = __float_as_uint(__uint_as_float(o8[i]) * alpha);
          tmem_st_x8(tm_o + lane_off + c, o8);
        }
      }
      pb = sb;
      if (!has_next) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full[sb]);
      }
      sb = nb;
      sph = nph;
    }

    // ---- epilogue: O / l -> bf16 -> global ------------------------------------------------------
    mbar_wait(&bars.o_done, 0);
    tc_fence_after();
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = P.o + ((long long)b * P.n_q + row) * P.ld_o + (long long)h * P.dv;
    for (int c = 0; c < P.dv; c += 8) {
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

template <int kStages, int kEmu, int kSBuf, bool kLateArrive>
int launch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mk2,
           const CUtensorMap& mv2, const Params& P, int batch, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)P.kb * kBM * 128 + (size_t)kStages * (P.kb + P.kbv) * kBN * 128;
  static size_t attr = 0;          // one process drives one GPU (vf_capi.cu: check_device binds the library to it)
  if (smem > attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(attn_stream_kernel<kStages, kEmu, kSBuf, kLateArrive>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  dim3 grid((P.n_q + kBM - 1) / kBM, batch * P.heads);
  attn_stream_kernel<kStages, kEmu, kSBuf, kLateArrive><<<grid, kThreads, smem, st>>>(mq, mk, mv, mk2, mv2, P);
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
  // d_head <= 128: one CTA per (query tile, batch, head).  Wider heads (multiples of 128 up to 512): Q and K keep the full
  // head for the scores, V and O are cut into vgroup column slices of 128 that the grid walks like heads.
  const bool wide = d > 128;
  if (wide && (d % 128 || d > 512)) return fail("vf_attn_fwd(bf16): d_head=%d > 192 must be a multiple of 128, <= 512", d);
  const int vgroup = wide ? d / 128 : 1;
  const int dv = wide ? 128 : d;
  Params P;
  P.o = reinterpret_cast<__nv_bfloat16*>(o);
  P.ld_o = ld_o;
  P.heads = heads * vgroup; P.n_q = n_q; P.n_kv = n_kv; P.n_kv2 = has2 ? n_kv2 : 0;
  P.d = d; P.d_pad = (d + 15) / 16 * 16; P.kb = (d + 63) / 64;
  P.dv = dv; P.dv_pad = (dv + 15) / 16 * 16; P.kbv = (dv + 63) / 64; P.vgroup = vgroup;
  P.scale_log2 = scale * 1.4426950408889634f;
  if ((long long)batch * P.heads > 65535) return fail("vf_attn_fwd: batch*heads*slices=%lld exceeds 65535", (long long)batch * P.heads);
  CUtensorMap mq, mk, mv, mk2, mv2;
  if (int rc = attn_make_map(&mq, q, batch, heads, n_q, d, ld_q, kBM)) return rc;
  if (int rc = attn_make_map(&mk, k, batch, heads, n_kv, d, ld_k, kBN)) return rc;
  if (int rc = attn_make_map(&mv, v, batch, P.heads, n_kv, dv, ld_v, kBN)) return rc;
  if (has2) {
    if (int rc = attn_make_map(&mk2, k2, batch, heads, n_kv2, d, ld_k2, kBN)) return rc;
    if (int rc = attn_make_map(&mv2, v2, batch, P.heads, n_kv2, dv, ld_v2, kBN)) return rc;
  } else {
    mk2 = mk;
    mv2 = mv;
  }
  // wide heads: Q alone is d/64 x 16 KB (128 KB at d = 512), so K and V get ONE stage each (<= 80 KB): 209 KB per CTA, one
  // CTA per SM.  A once-per-clip operator (first-stage decode / encode), not a per-step one.
  if (wide) return launch<1, 0, 2, false>(mq, mk, mv, mk2, mv2, P, batch, st);
  // VF_ATTN_SBUF = 2 | 3 (default 3 where it fits: d_pad <= 64), VF_ATTN_LATE = 0 | 1 (default 0): tuning knobs
  static int sbuf = -1, late = -1;
  if (sbuf < 0) { const char* e = getenv("VF_ATTN_SBUF"); sbuf = e ? atoi(e) : 3; }
  if (late < 0) { const char* e = getenv("VF_ATTN_LATE"); late = e ? atoi(e) : 0; }
  if (P.kb == 1) {                       // 16 KB Q + 3 x 16 KB K/V: 65 KB per CTA
    if (sbuf == 3 && P.d_pad <= 64) {
      if (late) {
        if (emu == 1) return launch<3, 1, 3, true>(mq, mk, mv, mk2, mv2, P, batch, st);
        return launch<3, 0, 3, true>(mq, mk, mv, mk2, mv2, P, batch, st);
      }
      switch (emu) {
        case 1: return launch<3, 1, 3, false>(mq, mk, mv, mk2, mv2, P, batch, st);
        case 2: return launch<3, 2, 3, false>(mq, mk, mv, mk2, mv2, P, batch, st);
        default: return launch<3, 0, 3, false>(mq, mk, mv, mk2, mv2, P, batch, st);
      }
    }
    if (late) return launch<3, 0, 2, true>(mq, mk, mv, mk2, mv2, P, batch, st);
    return launch<3, 0, 2, false>(mq, mk, mv, mk2, mv2, P, batch, st);
  }
  if (late) return launch<2, 0, 2, true>(mq, mk, mv, mk2, mv2, P, batch, st);      // kb = 2: 32 KB Q + 2 x 32 KB K/V: 97 KB per CTA
  return launch<2, 0, 2, false>(mq, mk, mv, mk2, mv2, P, batch, st);
}

}  // namespace vf
