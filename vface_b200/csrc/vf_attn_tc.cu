// Fused attention core on 5th-generation tensor cores (dtype VF_BF16 of vf_attn_fwd).
//
//   o = softmax(q k^T * scale) v   per (batch, head), q/k/v/o in the reference's native
//   (batch, n, heads*d) layout (ldm/models/pnp_utils.py:270-286, ldm/modules/attention.py:203-220),
//   optional second K/V segment concatenated along the key axis (injected target-frame K/V).
//
// Design (sm_100a):
//   * one CTA = one 128-row query tile of one (batch, head); 8 warps in two warpgroups:
//       warpgroup 0  (warps 0-3) softmax / correction / epilogue, one thread per query row (TMEM lane),
//                    with the registers of warpgroup 1 added (setmaxnreg.inc)
//       warpgroup 1  warp 4 = TMA producer (cp.async.bulk.tensor 4-D boxes, 128B swizzle, zero fill of
//                             the head-dim padding d -> 64k and of ragged row tails),
//                    warp 7 = tcgen05.mma issuer + TMEM allocator -- the highest warp id, so the scheduler
//                             serves it first (one elected lane issues, the loop is warp-uniform),
//                    warps 5,6 idle; the group gives its registers away (setmaxnreg.dec)
//   * S = Q K^T      : tcgen05.mma  SS, M=128, N=BN, K=16 x ceil(d/16), fp32 accumulator in TMEM
//   * P (bf16)       : written back over the first BN/2 columns of S by the softmax threads
//                      (tcgen05.st) once S is in registers -- never to smem/HBM
//   * O += P V       : tcgen05.mma  TS (A = P from TMEM), B = V tile in MN-major 128B-swizzled smem
//   * the tensor pipe executes in issue order, so PV_j followed by QK_{j+1} may reuse the S/P columns
//   * online softmax in the exp2 domain with lazy rescaling of O (only when the running max moves
//     by more than 2^8), so the correction pass is rare; FMNMX3 / FFMA2 / FADD2 keep the per-element
//     instruction count at ~3 so that the kernel is bound by MUFU.EX2, not by issue slots
//   * small CTAs (BN = 64: 48 KB smem, 128 TMEM columns at d <= 64) so that FOUR CTAs are resident per
//     SM: 4 softmax warps per SM sub-partition interleave their MUFU, TMEM-load and barrier phases.
//   * the head dimension is NOT padded in HBM: the TMA box is 64 elements wide over a tensor-map
//     dimension of extent d, out-of-bounds columns are zero-filled in shared memory.
//
// TMEM columns: [0,BN) S fp32, aliased by P bf16x2 in [0,BN/2) | [BN, BN+d_pad) O fp32.
#include "vf_attn.cuh"
#include "vf_sm100.cuh"

#include <cuda.h>
#include <cmath>
#include <cstdlib>

namespace vf {

using namespace sm100;

constexpr int kTcThreads = 256;
constexpr int kTmaWarp = 4;         // warps 0..3: softmax (one per TMEM lane quarter); 4: TMA producer; 7: MMA issuer
constexpr int kMmaWarp = 7;         //   (PV issuer; with aliased P the issuer of both chains)
constexpr int kQkWarp = 6;          // 6: QK issuer in the split-P modes
constexpr int kBM = 128;            // query rows per CTA
constexpr float kRescaleThreshold = 8.0f;   // log2 units
constexpr float kLagGuard = 64.0f;          // log2 units a row may exceed its lagged reference before the exact path

struct AttnTcParams {
  __nv_bfloat16* o;
  long long ld_o;
  int heads, n_q, n_kv, n_kv2, d, d_pad, kb;   // kb = ceil(d / 64) 64-wide head-dim blocks
  float scale_log2;                            // scale * log2(e)
  int spin;                                    // tuning knob: poll instead of suspending on the softmax-side barriers
  int issue;                                   // MMA issue order in the split-P modes (see the issuer section)
  int one;                                     // 1: trip count of the single-pass loop of the kFence variant
};

constexpr int kMaxStages = 4;

// Optional phase timing of the softmax warps (build with -DVF_ATTN_TRACE; never in the product build):
// per-phase clock64 totals of warp 4 of every CTA, summed into g_attn_trace and read with vf_attn_trace_read.
#ifdef VF_ATTN_TRACE
__device__ unsigned long long g_attn_trace[48];
__device__ unsigned long long g_attn_events[8][16];
#define VF_EV(j, e) do { if (blockIdx.x == 5 && blockIdx.y == 100 && (j) >= 20 && (j) < 28 && lane == 0) g_attn_events[(j) - 20][e] = clock64(); } while (0)
#define VF_TR_DECL unsigned long long tr_t = clock64(), tr_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define VF_TR(i) do { const unsigned long long n_ = clock64(); tr_acc[i] += n_ - tr_t; tr_t = n_; } while (0)
#else
#define VF_TR_DECL
#define VF_TR(i)
#define VF_EV(j, e)
#endif

struct __align__(8) TcBarriers {
  uint64_t q_full;
  uint64_t k_full[kMaxStages], k_empty[kMaxStages];
  uint64_t v_full[kMaxStages], v_empty[kMaxStages];
  uint64_t s_full, p_full, o_done;
  uint64_t s_free, p_empty;          // kSplitP only
  uint32_t tmem_base, tmem_base_p;
};

template <int kRegs> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" :: "n"(kRegs)); }
template <int kRegs> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" :: "n"(kRegs)); }

// kMinBlocks CTAs per SM; with 4 the launch-time register budget is 64 per thread and the softmax
// warpgroup is raised to 104 with the 40 that warpgroup 0 gives up.
// kEmu of every 4 score pairs take their exp2 on the FMA pipe (polynomial) instead of MUFU.
// kSplitP: P lives in its own 32-column TMEM allocation instead of aliasing S.  S_j is then free as soon
// as the softmax threads have loaded it into registers (s_free), so QK_{j+1} is issued while softmax_j is
// still in its exponentials and the softmax warps never wait for the tensor pipe; PV_j follows when P_j is
// complete (p_full) and releases P with its own commit (p_empty).  160 TMEM columns -> three CTAs per SM.
// kEarly (split-P modes): the two barrier waits of a tile (s_full of the NEXT tile, p_empty of the previous one) are
// probed with a non-blocking mbarrier.test_wait in the MIDDLE of the exponential section.  Both barriers complete
// hundreds of cycles before the softmax warps ask (profiles/r1_attn_trace_completion.txt); what the blocking form
// costs is the round trip of the probe itself through the instruction queue the MUFU stream of the other warps
// keeps full (400 + 110 cycles of a 2150-cycle tile).  Issued early, that round trip runs under this warp's own
// exponentials; the blocking wait remains as the fallback when a probe comes back negative.  kEarly = the element
// index of the section at which the probes are issued (S_{j+1} completes ~60 % into the section, PV_{j-1} ~50 %).
template <int BN, int kTmemCols, int kMinBlocks, int kEmu, int kPMode, int kStages, bool kLagMax, int kEarly = 0, bool kFence = false>
__global__ void __launch_bounds__(kTcThreads, kMinBlocks)
attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
               const __grid_constant__ CUtensorMap map_v, const __grid_constant__ CUtensorMap map_k2,
               const __grid_constant__ CUtensorMap map_v2, const AttnTcParams P) {
  // kPMode: 0 = P aliases S (128 columns, 4 CTAs/SM, QK_{j+1} behind PV_j);
  //         1 = P in its own BN/2-column allocation (BN = 64: 128 + 32 columns, 3 CTAs/SM);
  //         2 = P inside the main allocation behind O (BN = 48, d_pad <= 48: S 48 | O 48 | P 24 = 120 of 128
  //             columns, 4 CTAs/SM with the early QK_{j+1} of mode 1).
  constexpr bool kSplitP = kPMode != 0;
  static_assert(BN % 16 == 0 && (BN == 64 || BN == 48), "key tile of 48 or 64 rows");
  constexpr int kOCols = kTmemCols == 128 ? 48 : 160;      // O columns reserved in front of P (mode 2)
  static_assert(kPMode != 2 || BN + kOCols + BN / 2 <= kTmemCols, "S | O | P must fit the allocation");
  extern __shared__ unsigned char smem_dyn[];
  __shared__ TcBarriers bars;

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
  const uint32_t kv_block_bytes = BN * 128;
  const uint32_t q_bytes = P.kb * q_block_bytes;
  const uint32_t kv_bytes = P.kb * kv_block_bytes;
  unsigned char* sQ = tiles;
  unsigned char* sK = sQ + q_bytes;                 // kStages stages
  unsigned char* sV = sK + kStages * kv_bytes;      // kStages stages

  const int t1 = (P.n_kv + BN - 1) / BN;
  const int t2 = (P.n_kv2 + BN - 1) / BN;
  const int n_tiles = t1 + t2;

  if (threadIdx.x == 0) {
    mbar_init(&bars.q_full, 1);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&bars.k_full[s], 1);
      mbar_init(&bars.k_empty[s], 1);
      mbar_init(&bars.v_full[s], 1);
      mbar_init(&bars.v_empty[s], 1);
    }
    mbar_init(&bars.s_full, 1);
    mbar_init(&bars.p_full, 4);
    mbar_init(&bars.o_done, 1);
    mbar_init(&bars.s_free, 4);
    mbar_init(&bars.p_empty, 1);
    fence_barrier_init();
  }
  if (warp == kMmaWarp) {
    if (kPMode == 1) {
      tmem_alloc_only<kTmemCols>(&bars.tmem_base);
      tmem_alloc_only<BN / 2>(&bars.tmem_base_p);
      tmem_relinquish();
    } else {
      tmem_alloc<kTmemCols>(&bars.tmem_base);
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars.tmem_base;
  const uint32_t tm_s = tmem;          // S (fp32, BN columns); P (bf16x2) overwrites its first BN/2 columns
  const uint32_t tm_p = kPMode == 1 ? bars.tmem_base_p : kPMode == 2 ? tmem + BN + kOCols : tmem;
  const uint32_t tm_o = tmem + BN;

  if (warp >= 4) {
    if (kMinBlocks >= 3) reg_dec<24>();
    if (warp == kTmaWarp) {
      // =========================== TMA producer ==================================================
      if (lane == 0) {
        tma_prefetch_desc(&map_q);
        tma_prefetch_desc(&map_k);
        tma_prefetch_desc(&map_v);
        mbar_arrive_expect_tx(&bars.q_full, q_bytes);
        for (int kb = 0; kb < P.kb; ++kb)
          tma_load_4d(sQ + kb * q_block_bytes, &map_q, &bars.q_full, kb * 64, h, q_tile * kBM, b);
        for (int j = 0; j < n_tiles; ++j) {
          const int st = j % kStages;
          const uint32_t use = (uint32_t)(j / kStages);
          const bool seg2 = j >= t1;
          const int row0 = (seg2 ? j - t1 : j) * BN;
          const CUtensorMap* mk = seg2 ? &map_k2 : &map_k;
          const CUtensorMap* mv = seg2 ? &map_v2 : &map_v;
          mbar_wait(&bars.k_empty[st], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.k_full[st], kv_bytes);
          for (int kb = 0; kb < P.kb; ++kb)
            tma_load_4d(sK + st * kv_bytes + kb * kv_block_bytes, mk, &bars.k_full[st], kb * 64, h, row0, b);
          mbar_wait(&bars.v_empty[st], (use & 1) ^ 1);
          mbar_arrive_expect_tx(&bars.v_full[st], kv_bytes);
          for (int kb = 0; kb < P.kb; ++kb)
            tma_load_4d(sV + st * kv_bytes + kb * kv_block_bytes, mv, &bars.v_full[st], kb * 64, h, row0, b);
        }
      }
    } else if (warp == kMmaWarp || (kSplitP && warp == kQkWarp && P.issue == 1)) {
      // =========================== MMA issuer(s) =================================================
      // The issuing warp is the HIGHEST-numbered warp of the CTA (the scheduler arbitrates highest warp id
      // first): a tile needs seven small MMAs and, measured with clock64, a lowest-priority single thread
      // that competes with three busy softmax warps for issue slots needed ~200 cycles per tcgen05.mma.
      // The loop is warp-uniform (all lanes wait on the barriers, one elected lane issues), so the
      // descriptors are computed on the uniform datapath instead of moving through R2UR in front of every MMA.
      //
      // Issue order in the split-P modes (P.issue, VF_ATTN_ISSUE):
      //   0  one issuer, chains back to back:  wait s_free(j) -> QK_{j+1} (k_steps MMAs) ; wait p_full(j) -> PV_j
      //   1  (default) two issuers: warp 6 the QK chain, warp 7 the PV chain, each with its own commit stream
      //      (+1..3 %: neither chain queues behind the other's barrier wait).
      // Measured and dropped: one issuer alternating QK_{j+1} / PV_{j-1} MMAs -- 2x SLOWER in the kernel although
      // experiments/mma_mix.cu shows the pipe itself does not care about the order (35 clk per MMA SM-wide, 3 CTAs).
      {
        const uint32_t idesc_qk = make_idesc_bf16(kBM, BN, false);
        const uint32_t idesc_pv = make_idesc_bf16(kBM, P.d_pad, true);
        const int k_steps = P.d_pad / 16;
        const uint32_t q_addr = smem_u32(sQ);
        const uint32_t k_addr = smem_u32(sK);
        const uint32_t v_addr = smem_u32(sV);

        auto qk_mma = [&](int st, int s) {
          const uint32_t off_blk = (uint32_t)(s >> 2), off_in = (uint32_t)(s & 3) * 32u;
          const uint64_t da = make_smem_desc_sw128(q_addr + off_blk * q_block_bytes + off_in, 16, 1024);
          const uint64_t db = make_smem_desc_sw128(k_addr + st * kv_bytes + off_blk * kv_block_bytes + off_in, 16, 1024);
          if (elect_one()) mma_ss(tm_s, da, db, idesc_qk, s > 0);
        };
        auto pv_mma = [&](int st, int s, bool acc) {
          // B = V tile, MN-major: 64 head-dim elements contiguous (128 B) per key row, 8-row groups
          // 1024 B apart (SBO), further 64-wide head-dim blocks kv_block_bytes apart (LBO).
          const uint64_t db = make_smem_desc_sw128(v_addr + st * kv_bytes + (uint32_t)s * 2048u, kv_block_bytes, 1024);
          if (elect_one()) mma_ts(tm_o, tm_p + (uint32_t)s * 8u, db, idesc_pv, acc);
        };
        auto issue_qk = [&](int j) {
          const int st = j % kStages;
          mbar_wait(&bars.k_full[st], (uint32_t)(j / kStages) & 1);
          tc_fence_after();
          for (int s = 0; s < k_steps; ++s) qk_mma(st, s);
          if (elect_one()) {
            tc_commit(&bars.k_empty[st]);
            tc_commit(&bars.s_full);
          }
        };
        auto issue_pv = [&](int j) {
          const int st = j % kStages;
#pragma unroll
          for (int s = 0; s < BN / 16; ++s) pv_mma(st, s, (j > 0) || (s > 0));
          if (elect_one()) tc_commit(&bars.v_empty[st]);
        };
        auto wait_soft = [&](uint64_t* bar, uint32_t parity) {
          if (P.spin & 2) mbar_wait_spin(bar, parity);
          else mbar_wait(bar, parity);
        };

        VF_TR_DECL;
        if (kSplitP && P.issue == 1 && warp == kQkWarp) {
          // ---- QK issuer of the two-issuer mode: S_{j+1} as soon as the softmax threads hold S_j ---------
          mbar_wait(&bars.q_full, 0);
          issue_qk(0);
          for (int j = 0; j + 1 < n_tiles; ++j) {
            wait_soft(&bars.s_free, (uint32_t)j & 1);            // S_j is in the softmax threads' registers
            tc_fence_after();
            VF_TR(0);                     // QK issuer: wait s_free
            VF_EV(j, 8);                  // s_free(j) seen
            issue_qk(j + 1);
            VF_EV(j, 9);                  // QK(j+1) issued
            VF_TR(1);                     // issue QK (incl. k_full wait)
#ifdef VF_ATTN_TRACE
            mbar_wait(&bars.s_full, (uint32_t)(j + 1) & 1);     // trace build only: when does S_{j+1} complete?
            VF_EV(j, 12);
#endif
          }
        } else {
          // ---- PV issuer of the two-issuer mode / the single sequential issuer ---------------------------
          const bool both = !(kSplitP && P.issue == 1);
          if (both) {
            mbar_wait(&bars.q_full, 0);
            issue_qk(0);
          }
          for (int j = 0; j < n_tiles; ++j) {
            const int st = j % kStages;
            if (kSplitP && both && j + 1 < n_tiles) {
              wait_soft(&bars.s_free, (uint32_t)j & 1);
              tc_fence_after();
              VF_TR(0);
              VF_EV(j, 8);
              issue_qk(j + 1);
              VF_EV(j, 9);
              VF_TR(1);
            }
            mbar_wait(&bars.v_full[st], (uint32_t)(j / kStages) & 1);
            wait_soft(&bars.p_full, (uint32_t)j & 1);              // P_j in TMEM, O rescaled if needed
            tc_fence_after();
            VF_TR(2);                       // wait v_full + p_full
            VF_EV(j, 10);                   // p_full(j) seen
            issue_pv(j);
            if (kSplitP) {
              if (elect_one()) {
                tc_commit(&bars.p_empty);
                if (j + 1 == n_tiles) tc_commit(&bars.o_done);
              }
              VF_TR(3);                     // issue PV
              VF_EV(j, 11);                 // PV(j) issued
#ifdef VF_ATTN_TRACE
              if (P.issue == 1) {           // trace build only: when does PV_j complete?
                mbar_wait(&bars.p_empty, (uint32_t)j & 1);
                VF_EV(j, 13);
              }
#endif
            } else {
              if (j + 1 < n_tiles) issue_qk(j + 1);          // in order behind PV_j: may overwrite P_j
              else if (elect_one()) tc_commit(&bars.o_done);
            }
          }
        }
#ifdef VF_ATTN_TRACE
        if (lane == 0) for (int i = 0; i < 4; ++i) atomicAdd(&g_attn_trace[10 + i], tr_acc[i]);
#endif
      }
    }
  } else {
    // =========================== softmax / correction / epilogue ================================
    if (kMinBlocks >= 4) reg_inc<104>();
    else if (kMinBlocks == 3) reg_inc<136>();
    const int quarter = warp;                            // warps 0..3: TMEM lane quarter this warp may touch
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int row = q_tile * kBM + quarter * 32 + lane;  // query row owned by this thread
    float m_ref = 0.0f, l = 0.0f;
    float lag_alpha = 1.0f, lag_m = 0.0f;          // kLagMax: rescale decided by the previous tile's row max
    bool lag_need = false;
    bool s_probe = false, p_probe = false;         // kEarly: results of the probes issued inside the exponential section
    VF_TR_DECL;

    for (int j = 0; j < n_tiles; ++j) {
      const bool seg2 = j >= t1;
      const int row0 = (seg2 ? j - t1 : j) * BN;
      const int valid = min(BN, (seg2 ? P.n_kv2 : P.n_kv) - row0);

      // S_j complete; the commit also covers PV_{j-1}, so P/O are ours again.
      if (!(kEarly && kSplitP && j > 0 && s_probe)) {        // kEarly: already seen complete by the probe of tile j-1
        if (P.spin) mbar_wait_spin(&bars.s_full, (uint32_t)j & 1);
        else mbar_wait(&bars.s_full, (uint32_t)j & 1);
      }
      tc_fence_after();
      VF_TR(j == 0 ? 0 : 1);            // 0: prologue until S_0, 1: s_full waits
      if (warp == 0) VF_EV(j, 0);       // s_full(j) passed
      uint32_t sr[BN];
      tmem_ld_x32(tm_s + lane_off, *reinterpret_cast<uint32_t(*)[32]>(&sr[0]));
      if (BN == 64) tmem_ld_x32(tm_s + lane_off + 32, *reinterpret_cast<uint32_t(*)[32]>(&sr[32]));
      else tmem_ld_x16(tm_s + lane_off + 32, &sr[32]);
      if (kSplitP && j > 0) {
        // P_{j-1} hand-over, deferred to here so that its TMEM-store latency overlaps this tile's S load
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full);
      }
      tmem_wait_ld();
      if (kSplitP) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.s_free);
      }
      VF_TR(2);                         // S load (+ deferred P hand-over)
      if (warp == 0) VF_EV(j, 1);       // s_free(j) arrived (and p_full(j-1))

      if (valid < BN) {
#pragma unroll
        for (int i = 0; i < BN; ++i)
          if (i >= valid) sr[i] = 0xff800000u;   // -inf
      }
      // p = exp2(s * c - m_ref): packed FFMA2 for the affine part, MUFU.EX2 per element, packed FADD2
      // row sums in two independent chains, bf16x2 packing for the P operand; the row max (3-input FMNMX3 in
      // four independent chains) either up front or -- kLagMax -- inside the same instruction stream.
      const uint64_t c2 = pack2(P.scale_log2, P.scale_log2);
      uint64_t acc_a = 0ull, acc_b = 0ull;     // (+0.0f, +0.0f)
      uint32_t pk[BN / 2];
      float mx[4];
      auto row_max = [&]() {
#pragma unroll
        for (int t = 0; t < 4; ++t) mx[t] = __uint_as_float(sr[t]);
#pragma unroll
        for (int i = 0; i < BN; i += 8)
#pragma unroll
          for (int t = 0; t < 4; ++t)
            mx[t] = fmax3(mx[t], __uint_as_float(sr[i + 2 * t]), __uint_as_float(sr[i + 2 * t + 1]));
      };
      auto exps = [&](const bool with_max) {
        const uint64_t nm2 = pack2(-m_ref, -m_ref);
        acc_a = 0ull; acc_b = 0ull;
        if (with_max) {
#pragma unroll
          for (int t = 0; t < 4; ++t) mx[t] = __uint_as_float(sr[t]);
        }
#pragma unroll
        for (int i = 0; i < BN; i += 4) {
          if (kEarly && kSplitP && i == kEarly) {
            // probes for the two waits that follow this section; their round trip runs under the remaining MUFUs
            if (j + 1 < n_tiles) s_probe = mbar_test(&bars.s_full, (uint32_t)(j + 1) & 1);
            if (j > 0) p_probe = mbar_test(&bars.p_empty, (uint32_t)(j - 1) & 1);
          }
          const uint64_t xa = ffma2(pack2(__uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1])), c2, nm2);
          const uint64_t xb = ffma2(pack2(__uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3])), c2, nm2);
          if (with_max) {
            mx[(i / 4) & 3] = fmax3(mx[(i / 4) & 3], __uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1]));
            mx[(i / 4 + 2) & 3] = fmax3(mx[(i / 4 + 2) & 3], __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
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
          pk[i / 2 + 0] = pack_bf16(p0, p1);
          pk[i / 2 + 1] = pack_bf16(p2, p3);
        }
      };
      float alpha = 1.0f;
      bool need = false;
      if (!kLagMax) {
        row_max();
        const float cand = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]) * P.scale_log2;
        if (j == 0) {
          m_ref = cand;
        } else if (cand > m_ref + kRescaleThreshold) {
          alpha = ex2_approx(m_ref - cand);
          m_ref = cand;
          need = true;
        }
        VF_TR(3);                         // row max, rescale decision
        if (kFence) {
          // kFence: the 64 MUFU.EX2 as one bare stream, their consumers (row sums, bf16 packing) behind a single-pass loop
          // whose trip count ptxas does not know (a basic-block boundary).  Inside one block ptxas threads the consumers
          // between the MUFUs, where each stalls the in-order warp until its operands are back from the XU queue
          // (experiments/mufu_rate.cu: one warp per scheduler sustains 15.7 MUFU/clk/SM bare, 10.6 with them threaded in).
          const uint64_t nm2 = pack2(-m_ref, -m_ref);
          float ex[BN];
#pragma unroll
          for (int i = 0; i < BN; i += 2) {
            const uint64_t x2 = ffma2(pack2(__uint_as_float(sr[i + 0]), __uint_as_float(sr[i + 1])), c2, nm2);
            unpack2(x2, ex[i], ex[i + 1]);
          }
#pragma unroll
          for (int i = 0; i < BN; ++i) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(ex[i]) : "f"(ex[i]));
          acc_a = 0ull; acc_b = 0ull;
#pragma unroll
          for (int i = 0; i < BN / 2; ++i) pk[i] = 0u;
#pragma unroll 1
          for (int once = 0; once < P.one; ++once) __syncwarp();       // a block of its own between the stream and its consumers
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
        } else {
          exps(false);
        }
      } else {
        // Lagged reference: the exponentials of tile j use the reference decided from tiles < j (lag_alpha /
        // lag_need carry the O, l rescale that decision implies), while this tile's row max is computed in the
        // shadow of the MUFU stream and only decides the reference of tile j+1.  P_j may then exceed 1 -- by at
        // most 2^kLagGuard, harmless in bf16/fp32 -- and a row that jumps further (or tile 0, which has no
        // reference yet) takes the exact path: max first, exponentials again.  The result is the same softmax.
        alpha = lag_alpha;
        need = lag_need;
        bool redo = j == 0;
        if (!redo) {
          exps(true);
        } else {
          row_max();
        }
        const float cand = fmaxf(fmax3(mx[0], mx[1], mx[2]), mx[3]) * P.scale_log2;
        if (!redo) redo = __any_sync(0xffffffffu, cand > m_ref + kLagGuard);
        if (redo) {
          if (j == 0) {
            m_ref = cand;
          } else if (cand > m_ref + kRescaleThreshold) {
            alpha *= ex2_approx(m_ref - cand);
            m_ref = cand;
            need = true;
          }
          exps(false);
        }
        lag_alpha = 1.0f;
        lag_need = false;
        if (cand > m_ref + kRescaleThreshold) {     // decided now, applied to tile j+1 (its exps, then O and l)
          lag_alpha = ex2_approx(m_ref - cand);
          lag_m = cand;
          lag_need = true;
        }
      }
      VF_TR(4);                         // exponentials
      if (warp == 0) VF_EV(j, 2);       // exps done
      if (kSplitP && j > 0) {              // PV_{j-1} still reads P (and writes O) until its commit
        if (!(kEarly && p_probe)) {
          if (P.spin) mbar_wait_spin(&bars.p_empty, (uint32_t)(j - 1) & 1);
          else mbar_wait(&bars.p_empty, (uint32_t)(j - 1) & 1);
        }
        tc_fence_after();
      }
      VF_TR(5);                         // p_empty wait
      if (BN == 64) {
        tmem_st_x32(tm_p + lane_off, *reinterpret_cast<const uint32_t(*)[32]>(&pk[0]));
      } else {
        tmem_st_x16(tm_p + lane_off, &pk[0]);
        tmem_st_x8(tm_p + lane_off + 16, *reinterpret_cast<const uint32_t(*)[8]>(&pk[16]));
      }
      float sa0, sa1, sb0, sb1;
      unpack2(acc_a, sa0, sa1);
      unpack2(acc_b, sb0, sb1);
      l = l * alpha + ((sa0 + sa1) + (sb0 + sb1));
      if (kLagMax && lag_need) m_ref = lag_m;      // reference of the next tile
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
      VF_TR(6);                         // P store, row sums, rare O rescale
      if (warp == 0) VF_EV(j, 3);       // tile done
      if (!kSplitP || j + 1 == n_tiles) {
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars.p_full);
      }
    }

    VF_TR(6);                           // P store, row sums (tail of the last tile)
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
#ifdef VF_ATTN_TRACE
    VF_TR(7);                           // epilogue
    if (warp == 0 && lane == 0) {
      for (int i = 0; i < 8; ++i) atomicAdd(&g_attn_trace[i], tr_acc[i]);
      atomicAdd(&g_attn_trace[8], 1ull);
      atomicAdd(&g_attn_trace[9], (unsigned long long)n_tiles);
    }
    if (lane == 0) for (int i = 0; i < 8; ++i) atomicAdd(&g_attn_trace[16 + warp * 8 + i], tr_acc[i]);   // per quarter
#endif
  }

  __syncthreads();
  if (warp == kMmaWarp) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem);
    if (kPMode == 1) tmem_dealloc<BN / 2>(bars.tmem_base_p);
  }
}

// ---- host side --------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// (batch, n, heads*d) bf16 with row stride ld  ->  4-D map {d, heads, n, batch}, box {64, 1, rows, 1}.
int attn_make_map(CUtensorMap* m, const void* base, int batch, int heads, int n, int d, long long ld, int box_rows) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return fail("vf_attn_fwd: cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)heads, (cuuint64_t)n, (cuuint64_t)batch};
  cuuint64_t strides[3] = {(cuuint64_t)d * 2, (cuuint64_t)ld * 2, (cuuint64_t)ld * 2 * (cuuint64_t)n};
  cuuint32_t box[4] = {64, 1, (cuuint32_t)box_rows, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail("vf_attn_fwd: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

template <int BN, int kTmemCols, int kMinBlocks, int kEmu, int kPMode = 0, int kStages = 2, bool kLagMax = false, int kEarly = 0, bool kFence = false>
static int launch_tc(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, const CUtensorMap& mk2,
                     const CUtensorMap& mv2, const AttnTcParams& P, int batch, cudaStream_t st) {
  const size_t smem = 1024 + (size_t)P.kb * (kBM * 128 + 2 * kStages * BN * 128);
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(attn_tc_kernel<BN, kTmemCols, kMinBlocks, kEmu, kPMode, kStages, kLagMax, kEarly, kFence>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr = true;
  }
  dim3 grid((P.n_q + kBM - 1) / kBM, batch * P.heads);
  attn_tc_kernel<BN, kTmemCols, kMinBlocks, kEmu, kPMode, kStages, kLagMax, kEarly, kFence><<<grid, kTcThreads, smem, st>>>(mq, mk, mv, mk2, mv2, P);
  return check_cuda(cudaGetLastError(), "attn_tc_kernel launch");
}

int launch_attn_tc(const void* q, const void* k, const void* v, void* o, int batch, int heads, int n_q, int n_kv,
                   int d, long long ld_q, long long ld_k, long long ld_v, long long ld_o, float scale,
                   const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2, cudaStream_t st) {
  if (d % 8 || d < 8 || (d > 192 && (d % 128 || d > 512)))
    return fail("vf_attn_fwd(bf16): d_head=%d must be a multiple of 8 in [8, 192], or 256 / 384 / 512", d);
  const long long lds[4] = {ld_q, ld_k, ld_v, ld_o};
  for (long long ld : lds)
    if (ld % 8 || ld < (long long)heads * d) return fail("vf_attn_fwd(bf16): row strides must be multiples of 8 elements and >= heads*d");
  const void* ptrs[4] = {q, k, v, o};
  for (const void* p : ptrs)
    if (reinterpret_cast<uintptr_t>(p) & 15) return fail("vf_attn_fwd(bf16): pointers must be 16-byte aligned");
  const bool has2 = k2 != nullptr && n_kv2 > 0;
  if (has2) {
    if (!v2) return fail("vf_attn_fwd: k2 given without v2");
    if (ld_k2 % 8 || ld_v2 % 8 || ld_k2 < (long long)heads * d || ld_v2 < (long long)heads * d)
      return fail("vf_attn_fwd(bf16): bad k2/v2 row strides");
    if ((reinterpret_cast<uintptr_t>(k2) & 15) || (reinterpret_cast<uintptr_t>(v2) & 15))
      return fail("vf_attn_fwd(bf16): k2/v2 must be 16-byte aligned");
  }
  AttnTcParams P;
  P.o = reinterpret_cast<__nv_bfloat16*>(o);
  P.ld_o = ld_o;
  P.heads = heads; P.n_q = n_q; P.n_kv = n_kv; P.n_kv2 = has2 ? n_kv2 : 0;
  P.d = d; P.d_pad = (d + 15) / 16 * 16; P.kb = (d + 63) / 64;
  P.scale_log2 = scale * 1.4426950408889634f;
  {
    static int spin = -1;
    if (spin < 0) { const char* e = getenv("VF_ATTN_SPIN"); spin = e ? atoi(e) : 0; }
    P.spin = spin;
    static int issue = -1;
    if (issue < 0) { const char* e = getenv("VF_ATTN_ISSUE"); issue = e ? atoi(e) : 1; }
    P.issue = issue;
    P.one = 1;
  }
  static int emu = -1;      // tuning knob: VF_ATTN_EMU = 0..3 pairs of every 4 on the FMA pipe
  if (emu < 0) {
    const char* e = getenv("VF_ATTN_EMU");
    emu = e ? atoi(e) : 0;
    if (emu < 0 || emu > 3) emu = 0;
  }
  // VF_ATTN_SPLITP: 1 (default) P in its own TMEM allocation, key tiles of 64, 3 CTAs/SM; 3 = key tiles of 48 with
  // P behind O in one 128-column allocation, 4 CTAs/SM (d_pad <= 48); 0 = aliased P, 4 CTAs/SM; 2 = mode 1 with a
  // 3-stage K/V ring.
  static int split = -1;
  if (split < 0) {
    const char* e = getenv("VF_ATTN_SPLITP");
    split = e ? atoi(e) : 1;
  }
  // VF_ATTN_LAGMAX: 1 = row max of tile j in the shadow of its exponentials (lagged reference); 0 (default) = max first.
  // Measured equal within 1 % (gpurun_out/lagmax_ab.log): the row max is not on the kernel's critical resource.
  static int lag = -1;
  if (lag < 0) {
    const char* e = getenv("VF_ATTN_LAGMAX");
    lag = e ? atoi(e) : 0;
  }
  // VF_ATTN_STREAM: 0 (default) = the arrangements below for d_head <= 192, and the streamed-softmax arrangement of
  // vf_attn_stream.cu only for wide heads (d_head 256 / 384 / 512, which nothing else here can hold); 1 = streamed
  // arrangement also for d_head <= 64 (S triple-buffered in TMEM, two CTAs/SM); 2 = also for d_head <= 128.  Measured at
  // N = 4096, 8 x d40, 96 frame-branches (profiles/r2_attn_stream_ab.txt): 491-499 TF/s for the default against 428-451
  // for the streamed one -- with 8 softmax warps per SM instead of 12 its XU pipe sits at 60 %.
  static int stream_mode = -1;
  if (stream_mode < 0) {
    const char* e = getenv("VF_ATTN_STREAM");
    stream_mode = e ? atoi(e) : 0;
  }
  // VF_ATTN_PP: round-robin arrangement (vf_attn_pp.cu) for d_head <= 64 and at least 3 query tiles -- 1 / 2: one / two
  // softmax warps of a scheduler in their exponentials at a time, 3: no ordering; 0 = the arrangements below.
  static int pp_mode = -1;
  if (pp_mode < 0) { const char* e = getenv("VF_ATTN_PP"); pp_mode = e ? atoi(e) : 0; }
  if (pp_mode && P.d_pad <= 64 && n_q >= 3 * kBM)
    return launch_attn_pp(q, k, v, o, batch, heads, n_q, n_kv, d, ld_q, ld_k, ld_v, ld_o, scale, k2, v2, n_kv2, ld_k2, ld_v2,
                          pp_mode >= 3 ? 0 : pp_mode, st);
  if (d > 192 || (stream_mode && P.d_pad <= (stream_mode >= 2 ? 128 : 64)))
    return launch_attn_stream(q, k, v, o, batch, heads, n_q, n_kv, d, ld_q, ld_k, ld_v, ld_o, scale, k2, v2, n_kv2, ld_k2, ld_v2,
                              emu, st);
  const bool bn48 = split == 3 && P.d_pad <= 48;
  const int bn = bn48 ? 48 : 64;
  CUtensorMap mq, mk, mv, mk2, mv2;
  if (int rc = attn_make_map(&mq, q, batch, heads, n_q, d, ld_q, kBM)) return rc;
  if (int rc = attn_make_map(&mk, k, batch, heads, n_kv, d, ld_k, bn)) return rc;
  if (int rc = attn_make_map(&mv, v, batch, heads, n_kv, d, ld_v, bn)) return rc;
  if (has2) {
    if (int rc = attn_make_map(&mk2, k2, batch, heads, n_kv2, d, ld_k2, bn)) return rc;
    if (int rc = attn_make_map(&mv2, v2, batch, heads, n_kv2, d, ld_v2, bn)) return rc;
  } else {
    mk2 = mk;
    mv2 = mv;
  }
  if (bn48) {
    switch (emu) {
      case 1: return launch_tc<48, 128, 4, 1, 2>(mq, mk, mv, mk2, mv2, P, batch, st);
      default: return launch_tc<48, 128, 4, 0, 2>(mq, mk, mv, mk2, mv2, P, batch, st);
    }
  }
  // VF_ATTN_EARLY: early barrier probes (kEarly) -- 1: at element 48 of 64, 2: the same with the lagged row max, 3: with 25 %
  // of the exponentials on the FMA pipe, 4: both; 5 / 6: probes at element 32 / 56; 7: no probes, the exponentials as one bare
  // MUFU stream with their consumers fenced off (kFence); 0 = the blocking waits.
  static int early = -1;
  if (early < 0) { const char* e = getenv("VF_ATTN_EARLY"); early = e ? atoi(e) : 0; }
  if (P.d_pad <= 64 && split == 1 && early) {
    switch (early) {
      case 2: return launch_tc<64, 128, 3, 0, 1, 2, true, 48>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 3: return launch_tc<64, 128, 3, 1, 1, 2, false, 48>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 4: return launch_tc<64, 128, 3, 1, 1, 2, true, 48>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 5: return launch_tc<64, 128, 3, 0, 1, 2, false, 32>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 6: return launch_tc<64, 128, 3, 0, 1, 2, false, 56>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 7: return launch_tc<64, 128, 3, 0, 1, 2, false, 0, true>(mq, mk, mv, mk2, mv2, P, batch, st);     // bare MUFU stream
      default: return launch_tc<64, 128, 3, 0, 1, 2, false, 48>(mq, mk, mv, mk2, mv2, P, batch, st);
    }
  }
  if (split && P.d_pad > 64 && P.d_pad <= 160 && early) {
    if (early == 2 || early == 4) return launch_tc<64, 256, 2, 0, 2, 2, true, 48>(mq, mk, mv, mk2, mv2, P, batch, st);
    return launch_tc<64, 256, 2, 0, 2, 2, false, 48>(mq, mk, mv, mk2, mv2, P, batch, st);
  }
  if (P.d_pad <= 64 && split == 2) return launch_tc<64, 128, 3, 0, 1, 3>(mq, mk, mv, mk2, mv2, P, batch, st);   // 3-stage K/V ring
  if (P.d_pad <= 64 && split) {
    switch (emu) {
      case 1: return launch_tc<64, 128, 3, 1, 1>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 2: return launch_tc<64, 128, 3, 2, 1>(mq, mk, mv, mk2, mv2, P, batch, st);
      default:
        if (lag) return launch_tc<64, 128, 3, 0, 1, 2, true>(mq, mk, mv, mk2, mv2, P, batch, st);
        return launch_tc<64, 128, 3, 0, 1>(mq, mk, mv, mk2, mv2, P, batch, st);
    }
  }
  if (P.d_pad <= 64) {                                                                   // 48 KB smem: 4 CTAs / SM
    switch (emu) {
      case 0: return launch_tc<64, 128, 4, 0>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 2: return launch_tc<64, 128, 4, 2>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 3: return launch_tc<64, 128, 4, 3>(mq, mk, mv, mk2, mv2, P, batch, st);
      case 1: return launch_tc<64, 128, 4, 1>(mq, mk, mv, mk2, mv2, P, batch, st);
      default: return launch_tc<64, 128, 4, 0>(mq, mk, mv, mk2, mv2, P, batch, st);
    }
  }
  // d in (64, 192]: 256 TMEM columns, two CTAs/SM.  Up to d = 160 (the UNet's 80 and 160) P sits behind O in
  // the same allocation (S 64 | O 160 | P 32) so QK_{j+1} is issued early as in the d <= 64 kernel.
  if (split && P.d_pad <= 160) {
    if (lag) return launch_tc<64, 256, 2, 0, 2, 2, true>(mq, mk, mv, mk2, mv2, P, batch, st);
    return launch_tc<64, 256, 2, 0, 2>(mq, mk, mv, mk2, mv2, P, batch, st);
  }
  return launch_tc<64, 256, 2, 0>(mq, mk, mv, mk2, mv2, P, batch, st);
}

}  // namespace vf

#ifdef VF_ATTN_TRACE
extern "C" int vf_attn_trace_read(unsigned long long* out16, int reset) {
  if (reset == 2) return cudaMemcpyFromSymbol(out16, vf::g_attn_events, sizeof(unsigned long long) * 128) != cudaSuccess;
  if (cudaMemcpyFromSymbol(out16, vf::g_attn_trace, sizeof(unsigned long long) * 48) != cudaSuccess) return 1;
  if (reset) {
    unsigned long long z[48] = {0};
    if (cudaMemcpyToSymbol(vf::g_attn_trace, z, sizeof(z)) != cudaSuccess) return 1;
  }
  return 0;
}
#endif
