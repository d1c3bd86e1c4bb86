// Fused memory-bound glue kernels of the UNet between the GEMM-shaped ops:
//
//   vf_group_norm_nhwc   GroupNorm32 (+ per-(sample,channel) additive vector, + SiLU) on channels-last
//                        activations; replaces GroupNorm32 -> SiLU (openaimodel.py:201-205, :236-239,
//                        util.py:214-216) and the `h + emb_out` add of ResBlock._forward (:265-273).
//   vf_add_layer_norm    res = x + y + bias ; out = LayerNorm(res)   (attention.py:239-243)
//   vf_geglu             out = h[:, :k] * gelu(h[:, k:])              (attention.py:37-45)
//   vf_add_bias          out = a + b + bias[c]                         (residual adds, conv bias)
//
// All are HBM-bound: every activation is read/written exactly once per kernel with 16-byte vectors,
// statistics in fp32.  GroupNorm is two passes (statistics, apply): 2 reads + 1 write of x, against
// 5 passes + 2 layout conversions for the eager GroupNorm(NCHW)/SiLU/add chain it replaces.
#include "vf_common.cuh"
#include "vf_sm100.cuh"

#include <cooperative_groups.h>
#include <cstdlib>

namespace vf {

template <typename T> struct V16;   // 16-byte vector of T as floats
template <> struct V16<float> {
  static constexpr int E = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&v)[4]) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y); v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[4]) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
  static __device__ __forceinline__ void st(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  }
};
template <> struct V16<__nv_bfloat16> {
  static constexpr int E = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&v)[8]) {
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  }
  static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
    *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
};

// 16 raw bytes -> E/2 register pairs for the packed fp32 pipe (FFMA2 / FADD2), and back
template <typename T> struct P16;
template <> struct P16<float> {
  static __device__ __forceinline__ void unpack(const uint4& u, float2 (&v)[2]) {
    v[0] = make_float2(__uint_as_float(u.x), __uint_as_float(u.y));
    v[1] = make_float2(__uint_as_float(u.z), __uint_as_float(u.w));
  }
  static __device__ __forceinline__ uint4 pack(const float2 (&v)[2]) {
    return make_uint4(__float_as_uint(v[0].x), __float_as_uint(v[0].y), __float_as_uint(v[1].x), __float_as_uint(v[1].y));
  }
};
template <> struct P16<__nv_bfloat16> {
  static __device__ __forceinline__ void unpack(const uint4& u, float2 (&v)[4]) {
    v[0] = make_float2(bf16lo(u.x), bf16hi(u.x)); v[1] = make_float2(bf16lo(u.y), bf16hi(u.y));
    v[2] = make_float2(bf16lo(u.z), bf16hi(u.z)); v[3] = make_float2(bf16lo(u.w), bf16hi(u.w));
  }
  static __device__ __forceinline__ uint4 pack(const float2 (&v)[4]) {
    return make_uint4(pack_bf16(v[0].x, v[0].y), pack_bf16(v[1].x, v[1].y), pack_bf16(v[2].x, v[2].y), pack_bf16(v[3].x, v[3].y));
  }
};

// SiLU.  fp32 path: IEEE division (reference precision).  bf16 path: the output is rounded to 8 mantissa
// bits anyway, so the quotient goes through the fast reciprocal (MUFU.RCP) instead of the ~10-instruction
// IEEE division sequence.
template <typename T> __device__ __forceinline__ float silu_f(float t);
template <> __device__ __forceinline__ float silu_f<float>(float t) { return t / (1.0f + __expf(-t)); }
template <> __device__ __forceinline__ float silu_f<__nv_bfloat16>(float t) {
  // t sigmoid(t) = h + h tanh(h), h = t / 2: ONE MUFU (tanh.approx, |rel err| < 2^-10.9) instead of EX2 + RCP -- the
  // apply pass issued 16 MUFU per 16-byte chunk and showed mio_throttle stalls; the result is rounded to bf16.
  const float h = 0.5f * t;
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(h));
  return fmaf(h, th, h);
}

// ================================================================================================
// GroupNorm (NHWC)
// ================================================================================================
constexpr int kGnThreads = 256;
#ifndef VF_GN_MINB
#define VF_GN_MINB 1
#endif
constexpr int kGnMaxGroups = 32;

struct GnParams {
  const void* x; const void* add_nc; const void* gamma; const void* beta; void* y;
  const void* x2;       // second source (channels [c1, c)) of a channel-concatenated input, or null
  int c1;               // channels taken from x (== c when x2 is null)
  float* ws;            // [n][slabs][groups][2] partial (sum, sumsq)
  int n, hw, c, groups, slabs, rows_per_slab;
  float eps;
  int silu;
};

// thread -> (row lane, chunk) mapping shared by both passes
struct GnMap {
  int chunks, tpr, rpb, my_row, my_chunk;
  bool active;
  __device__ GnMap(int c, int elems) {
    chunks = c / elems;
    tpr = chunks < kGnThreads ? chunks : kGnThreads;     // threads per row
    rpb = kGnThreads / tpr;                               // rows per block iteration
    my_row = threadIdx.x / tpr;
    my_chunk = threadIdx.x - my_row * tpr;
    active = my_row < rpb;
  }
};

template <typename T, int CPT>    // CPT: chunks per thread along the channel axis (c/E <= 256*CPT)
__global__ void __launch_bounds__(kGnThreads, VF_GN_MINB)
gn_stats_kernel(const GnParams P) {
  constexpr int E = V16<T>::E;
  // per-(row lane, channel) partials, reduced to groups in a fixed order: bit-reproducible statistics
  extern __shared__ float s_part[];      // [2][rpb * c]
  const int n = blockIdx.y, slab = blockIdx.x;
  const GnMap M(P.c, E);
  float* s_psum = s_part;
  float* s_psq = s_part + M.rpb * P.c;
  const int cpg = P.c / P.groups;
  const int r0 = slab * P.rows_per_slab;
  const int r1 = min(P.hw, r0 + P.rows_per_slab);
  const int c2 = P.c - P.c1;
  const T* x = reinterpret_cast<const T*>(P.x) + (size_t)n * P.hw * P.c1;
  const T* x2 = P.x2 ? reinterpret_cast<const T*>(P.x2) + (size_t)n * P.hw * c2 : nullptr;
  const T* add = P.add_nc ? reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c : nullptr;
  if (M.active) {
    float sum[CPT][E], sq[CPT][E], av[CPT][E];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
#pragma unroll
      for (int j = 0; j < E; ++j) { sum[k][j] = 0.f; sq[k][j] = 0.f; av[k][j] = 0.f; }
      if (add && ch < P.c) V16<T>::ld(add + ch, av[k]);
    }
    // kRowUnroll rows per trip: that many independent 16-byte loads in flight per thread and chunk (one load
    // per trip left the SM far short of the ~44 KB it must keep in flight to cover HBM latency)
    constexpr int kRowUnroll = CPT == 1 ? 4 : 2;
    int r = r0 + M.my_row;
    for (; r + (kRowUnroll - 1) * M.rpb < r1; r += kRowUnroll * M.rpb) {
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int ch = (M.my_chunk + k * kGnThreads) * E;
        if (ch < P.c) {
          uint4 raw[kRowUnroll];
#pragma unroll
          for (int u = 0; u < kRowUnroll; ++u) {
            const size_t rr = (size_t)(r + u * M.rpb);
            raw[u] = ld_nc_v4(ch < P.c1 ? x + rr * P.c1 + ch : x2 + rr * c2 + (ch - P.c1));
          }
#pragma unroll
          for (int u = 0; u < kRowUnroll; ++u) {
            float v[E];
            V16<T>::unpack(raw[u], v);
#pragma unroll
            for (int j = 0; j < E; ++j) { const float t = v[j] + av[k][j]; sum[k][j] += t; sq[k][j] = fmaf(t, t, sq[k][j]); }
          }
        }
      }
    }
    for (; r < r1; r += M.rpb) {
#pragma unroll
      for (int k = 0; k < CPT; ++k) {
        const int ch = (M.my_chunk + k * kGnThreads) * E;
        if (ch < P.c) {
          float v[E];
          V16<T>::ld(ch < P.c1 ? x + (size_t)r * P.c1 + ch : x2 + (size_t)r * c2 + (ch - P.c1), v);
#pragma unroll
          for (int j = 0; j < E; ++j) { const float t = v[j] + av[k][j]; sum[k][j] += t; sq[k][j] = fmaf(t, t, sq[k][j]); }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
      if (ch < P.c) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          s_psum[M.my_row * P.c + ch + j] = sum[k][j];
          s_psq[M.my_row * P.c + ch + j] = sq[k][j];
        }
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < P.groups) {
    float gs = 0.f, gq = 0.f;
    for (int r = 0; r < M.rpb; ++r)
      for (int ch = threadIdx.x * cpg; ch < (threadIdx.x + 1) * cpg; ++ch) {
        gs += s_psum[r * P.c + ch];
        gq += s_psq[r * P.c + ch];
      }
    float* w = P.ws + (((size_t)n * P.slabs + slab) * P.groups + threadIdx.x) * 2;
    w[0] = gs;
    w[1] = gq;
  }
}

template <typename T, int CPT>
__global__ void __launch_bounds__(kGnThreads, VF_GN_MINB)
gn_apply_kernel(const GnParams P) {
  constexpr int E = V16<T>::E;
  __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
  const int n = blockIdx.y, slab = blockIdx.x;
  const int cpg = P.c / P.groups;
  if (threadIdx.x < P.groups) {
    double s = 0.0, q = 0.0;
    for (int i = 0; i < P.slabs; ++i) {
      const float* w = P.ws + (((size_t)n * P.slabs + i) * P.groups + threadIdx.x) * 2;
      s += (double)w[0];
      q += (double)w[1];
    }
    const double cnt = (double)P.hw * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    s_mean[threadIdx.x] = (float)mean;
    s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)P.eps));
  }
  __syncthreads();
  const GnMap M(P.c, E);
  if (!M.active) return;
  const int r0 = slab * P.rows_per_slab;
  const int r1 = min(P.hw, r0 + P.rows_per_slab);
  const int c2 = P.c - P.c1;
  const T* x = reinterpret_cast<const T*>(P.x) + (size_t)n * P.hw * P.c1;
  const T* x2 = P.x2 ? reinterpret_cast<const T*>(P.x2) + (size_t)n * P.hw * c2 : nullptr;
  T* y = reinterpret_cast<T*>(P.y) + (size_t)n * P.hw * P.c;
  const T* add = P.add_nc ? reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c : nullptr;
  const T* gamma = reinterpret_cast<const T*>(P.gamma);
  const T* beta = reinterpret_cast<const T*>(P.beta);
  float scale[CPT][E], shift[CPT][E];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int ch = (M.my_chunk + k * kGnThreads) * E;
    if (ch < P.c) {
      float g[E], b[E], a[E];
      V16<T>::ld(gamma + ch, g);
      V16<T>::ld(beta + ch, b);
#pragma unroll
      for (int j = 0; j < E; ++j) a[j] = 0.f;
      if (add) V16<T>::ld(add + ch, a);
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int grp = (ch + j) / cpg;
        scale[k][j] = s_rstd[grp] * g[j];
        shift[k][j] = fmaf(a[j] - s_mean[grp], scale[k][j], b[j]);   // (x + a - mean) * rstd * gamma + beta
      }
    }
  }
  constexpr int kRowUnroll = CPT == 1 ? 4 : 2;
  int r = r0 + M.my_row;
  for (; r + (kRowUnroll - 1) * M.rpb < r1; r += kRowUnroll * M.rpb) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
      if (ch < P.c) {
        uint4 raw[kRowUnroll];
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
          const size_t rr = (size_t)(r + u * M.rpb);
          raw[u] = ld_nc_v4(ch < P.c1 ? x + rr * P.c1 + ch : x2 + rr * c2 + (ch - P.c1));
        }
#pragma unroll
        for (int u = 0; u < kRowUnroll; ++u) {
          float v[E];
          V16<T>::unpack(raw[u], v);
#pragma unroll
          for (int j = 0; j < E; ++j) {
            float t = fmaf(v[j], scale[k][j], shift[k][j]);
            if (P.silu) t = silu_f<T>(t);
            v[j] = t;
          }
          st_na_v4(y + (size_t)(r + u * M.rpb) * P.c + ch, V16<T>::pack(v));
        }
      }
    }
  }
  for (; r < r1; r += M.rpb) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
      if (ch < P.c) {
        float v[E];
        V16<T>::ld(ch < P.c1 ? x + (size_t)r * P.c1 + ch : x2 + (size_t)r * c2 + (ch - P.c1), v);
#pragma unroll
        for (int j = 0; j < E; ++j) {
          float t = fmaf(v[j], scale[k][j], shift[k][j]);
          if (P.silu) t = silu_f<T>(t);
          v[j] = t;
        }
        V16<T>::st(y + (size_t)r * P.c + ch, v);
      }
    }
  }
}

static void gn_plan(int n, int hw, int* slabs, int* rows_per_slab) {
  // CTAs per SM the (slabs x samples) grid aims for.  Swept 4..24 on B200 (VF_GN_CTAS_PER_SM): 6 is the best at every
  // level of the UNet (64x64x320: 64 -> 73 % of HBM; cat(320+320): 62 -> 74 %): 4 leaves a partial second wave of
  // fat CTAs, 16+ pays per-CTA prologue and statistics reduction.
  static int per_sm = -1;
  if (per_sm < 0) { const char* e_ = getenv("VF_GN_CTAS_PER_SM"); per_sm = e_ ? atoi(e_) : 6; if (per_sm < 1) per_sm = 6; }
  int want = (per_sm * num_sms() + n - 1) / n;
  if (want < 1) want = 1;
  if (want > 64) want = 64;
  int rps = (hw + want - 1) / want;
  if (rps < 8) rps = hw < 8 ? hw : 8;
  *rows_per_slab = rps;
  *slabs = (hw + rps - 1) / rps;
}

// ---- fused GroupNorm: ONE persistent launch, the apply pass re-reads its slab out of L2 ----------------------
// The two launches above stream x from HBM twice (the tensor of a 96-sample step does not fit the 126 MB L2 between
// them) and pay two grid tails.  Here a persistent grid takes (sample, slab) work items off a ticket counter in
// sample-major order; a CTA reduces the statistics of its slab, publishes them, waits until every slab of ITS
// sample has arrived (a per-sample counter: the other slabs are held by CTAs that took their tickets at about the
// same time), and applies the normalisation to the SAME slab right away -- the second read of x hits L2 as long as
// (resident CTAs x slab bytes) stays well below the L2 size, which the slab plan below guarantees (~36 MB).
//   * tickets make the wait deadlock-free without assuming a block dispatch order: every lower ticket is held by a
//     CTA that is already running (it only waits on tickets lower than or next to its own);
//   * statistics stay bit-reproducible: per-slab partials in a fixed order, per-sample totals in a fixed order;
//   * the counters clean up after themselves (the last CTA to leave a sample / the grid resets them), live in
//     module-global memory and are rotated over kGnSyncSets launches.  Launches in stream order (the sampler's single
//     compute stream, CUDA-graph replays included) never share a live set; GroupNorm launches that run CONCURRENTLY on
//     different streams are only safe while fewer than kGnSyncSets of them are in flight (VF_GN_FUSED=0 has no such state).
//   * the block is tpr x rpb threads (chunks per row x rows per trip, rounded up to a warp) so no lane idles at
//     c = 1280 / 1920 / 2560, and every thread keeps kGnfUnroll 16-byte loads in flight.
constexpr int kGnfMaxThreads = 320;
constexpr int kGnSyncSets = 8;
constexpr int kGnfMaxN = 4096;
__device__ unsigned int g_gn_sync[kGnSyncSets][2 + 2 * kGnfMaxN];     // ticket, finished, then (arrived, departed) per sample

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

template <typename T, int U, int kMinB>     // U 16-byte loads in flight per thread, kMinB CTAs per SM
__global__ void __launch_bounds__(kGnfMaxThreads, kMinB)
gn_fused_kernel(const GnParams P, const int tpr, const int rpb, const int set) {
  constexpr int E = V16<T>::E;
  extern __shared__ float s_part[];                     // [2][rpb * c] per-(row lane, channel) partials
  __shared__ float s_mean[kGnMaxGroups], s_rstd[kGnMaxGroups];
  __shared__ double s_red[kGnfMaxThreads / 32][kGnMaxGroups][2];
  __shared__ int s_item;
  unsigned int* sync = g_gn_sync[set];
  const int tid = threadIdx.x;
  const int my_row = tid / tpr, my_chunk = tid - my_row * tpr;
  const bool active = my_row < rpb;
  const int ch = my_chunk * E;
  const int cpg = P.c / P.groups;
  const int c2 = P.c - P.c1;
  const int total = P.n * P.slabs;
  float* s_psum = s_part;
  float* s_psq = s_part + rpb * P.c;
  const bool from1 = ch < P.c1;
  const int ld = from1 ? P.c1 : c2;                     // row stride of the source this thread reads
  const int ch_src = from1 ? ch : ch - P.c1;

  for (;;) {
    if (tid == 0) s_item = (int)atomicAdd(&sync[0], 1u);
    __syncthreads();
    const int item = s_item;
    if (item >= total) break;
    const int n = item / P.slabs, slab = item - n * P.slabs;
    const int r0 = slab * P.rows_per_slab;
    const int r1 = min(P.hw, r0 + P.rows_per_slab);
    const T* src = (from1 ? reinterpret_cast<const T*>(P.x) : reinterpret_cast<const T*>(P.x2)) + (size_t)n * P.hw * ld + ch_src;
    T* dst = reinterpret_cast<T*>(P.y) + (size_t)n * P.hw * P.c + ch;
    // ---- phase 1: statistics of the slab ---------------------------------------------------------------
    // The additive vector a[c] stays out of the loop (and out of the registers the loads in flight need):
    // sum(x + a) = sum(x) + N a,  sum((x + a)^2) = sum(x^2) + a (2 sum(x) + N a).
    if (active) {
      float sum[E], sq[E];
#pragma unroll
      for (int j = 0; j < E; ++j) { sum[j] = 0.f; sq[j] = 0.f; }
      int r = r0 + my_row;
      const int my_rows = r < r1 ? (r1 - r + rpb - 1) / rpb : 0;
      // one running pointer, all U loads issued before the first use (address arithmetic per load would eat the
      // registers the loads in flight need and ptxas then interleaves use and issue: two loads in flight, not U)
      const char* p = reinterpret_cast<const char*>(src) + (size_t)r * ld * sizeof(T);
      const size_t stride = (size_t)rpb * ld * sizeof(T);
      for (; r + (U - 1) * rpb < r1; r += U * rpb) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { raw[u] = ld_nc_v4(p); p += stride; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float v[E];
          V16<T>::unpack(raw[u], v);
#pragma unroll
          for (int j = 0; j < E; ++j) { sum[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]); }
        }
      }
      for (; r < r1; r += rpb, p += stride) {             // ragged tail (the slab plan makes it rare)
        float v[E];
        V16<T>::unpack(ld_nc_v4(p), v);
#pragma unroll
        for (int j = 0; j < E; ++j) { sum[j] += v[j]; sq[j] = fmaf(v[j], v[j], sq[j]); }
      }
      if (P.add_nc) {
        float av[E];
        V16<T>::ld(reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c + ch, av);
        const float cnt = (float)my_rows;
#pragma unroll
        for (int j = 0; j < E; ++j) {
          sq[j] = fmaf(av[j], fmaf(cnt, av[j], 2.0f * sum[j]), sq[j]);
          sum[j] = fmaf(cnt, av[j], sum[j]);
        }
      }
#pragma unroll
      for (int j = 0; j < E; ++j) {
        s_psum[my_row * P.c + ch + j] = sum[j];
        s_psq[my_row * P.c + ch + j] = sq[j];
      }
    }
    __syncthreads();
    // eight lanes per group, channels strided over the lanes, row lanes in order, then a fixed shuffle tree
    for (int grp = tid >> 3; grp < P.groups; grp += blockDim.x >> 3) {       // blockDim is a multiple of 32
      const int sub = tid & 7;
      float gs = 0.f, gq = 0.f;
      for (int i = sub; i < cpg; i += 8)
        for (int rr = 0; rr < rpb; ++rr) {
          gs += s_psum[rr * P.c + grp * cpg + i];
          gq += s_psq[rr * P.c + grp * cpg + i];
        }
      const unsigned m8 = 0xffu << (tid & 24);           // the eight lanes of this group (they share grp)
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(m8, gs, o);
        gq += __shfl_xor_sync(m8, gq, o);
      }
      if (sub == 0) {
        float* w = P.ws + (((size_t)n * P.slabs + slab) * P.groups + grp) * 2;
        __stcg(w, gs);
        __stcg(w + 1, gq);
      }
    }
    __syncthreads();
    if (tid == 0) {
      __threadfence();                                   // the CTA's partials (ordered by the barrier) before the arrival
      atomicAdd(&sync[2 + 2 * n], 1u);
      while (ld_acquire_u32(&sync[2 + 2 * n]) < (unsigned)P.slabs) __nanosleep(64);
      // everyone who departs has passed the wait: the last one may clear the sample's counters for the next launch
      const unsigned d = atomicAdd(&sync[3 + 2 * n], 1u);
      if (d == (unsigned)P.slabs - 1u) { sync[2 + 2 * n] = 0u; sync[3 + 2 * n] = 0u; }
    }
    __syncthreads();

    // ---- totals of the sample (fixed order), mean / rstd per group -------------------------------------
    {
      const int grp = tid & 31, part = tid >> 5, parts = blockDim.x >> 5;
      if (grp < P.groups) {
        double s = 0.0, q = 0.0;
        for (int i = part; i < P.slabs; i += parts) {
          const float* w = P.ws + (((size_t)n * P.slabs + i) * P.groups + grp) * 2;
          s += (double)__ldcg(w);
          q += (double)__ldcg(w + 1);
        }
        s_red[part][grp][0] = s;
        s_red[part][grp][1] = q;
      }
      __syncthreads();
      if (tid < P.groups) {
        double s = 0.0, q = 0.0;
        for (int p = 0; p < parts; ++p) { s += s_red[p][tid][0]; q += s_red[p][tid][1]; }
        const double cnt = (double)P.hw * cpg;
        const double mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[tid] = (float)mean;
        s_rstd[tid] = (float)(1.0 / sqrt(var + (double)P.eps));
      }
      __syncthreads();
    }

    // ---- phase 2: apply to the same slab (x comes back out of L2) ------------------------------------------
    if (active) {
      float scale[E], shift[E], gam[E], bet[E];          // gamma / beta re-read per item (L1/L2 hits): 16 registers
      V16<T>::ld(reinterpret_cast<const T*>(P.gamma) + ch, gam);      // that phase 1 needs for loads in flight
      V16<T>::ld(reinterpret_cast<const T*>(P.beta) + ch, bet);
      float av[E];
#pragma unroll
      for (int j = 0; j < E; ++j) av[j] = 0.f;
      if (P.add_nc) V16<T>::ld(reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c + ch, av);
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int grp = (ch + j) / cpg;
        scale[j] = s_rstd[grp] * gam[j];
        shift[j] = fmaf(av[j] - s_mean[grp], scale[j], bet[j]);        // (x + a - mean) * rstd * gamma + beta
      }
      int r = r0 + my_row;
      const char* p = reinterpret_cast<const char*>(src) + (size_t)r * ld * sizeof(T);
      char* q = reinterpret_cast<char*>(dst) + (size_t)r * P.c * sizeof(T);
      const size_t stride = (size_t)rpb * ld * sizeof(T);
      const size_t qstride = (size_t)rpb * P.c * sizeof(T);
      for (; r + (U - 1) * rpb < r1; r += U * rpb) {
        uint4 raw[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { raw[u] = ld_nc_v4(p); p += stride; }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          float v[E];
          V16<T>::unpack(raw[u], v);
#pragma unroll
          for (int j = 0; j < E; ++j) {
            float t = fmaf(v[j], scale[j], shift[j]);
            if (P.silu) t = silu_f<T>(t);
            v[j] = t;
          }
          st_na_v4(q, V16<T>::pack(v));
          q += qstride;
        }
      }
      for (; r < r1; r += rpb, p += stride, q += qstride) {
        float v[E];
        V16<T>::unpack(ld_nc_v4(p), v);
#pragma unroll
        for (int j = 0; j < E; ++j) {
          float t = fmaf(v[j], scale[j], shift[j]);
          if (P.silu) t = silu_f<T>(t);
          v[j] = t;
        }
        st_na_v4(q, V16<T>::pack(v));
      }
    }
  }
  // the last CTA out re-arms the ticket counter (nobody takes a ticket after leaving the loop)
  if (tid == 0) {
    __threadfence();
    const unsigned f = atomicAdd(&sync[1], 1u);
    if (f == gridDim.x - 1u) { sync[0] = 0u; sync[1] = 0u; }
  }
}

// VF_GN_FUSED_CTAS: 2 (default) = two CTAs per SM with eight loads in flight per thread (96 registers);
// 3 = three CTAs per SM with four (64 registers: eight spill).
static int gnf_ctas_per_sm() {
  static int v = -1;
  if (v < 0) { const char* e_ = getenv("VF_GN_FUSED_CTAS"); v = e_ ? atoi(e_) : 2; if (v != 3) v = 2; }
  return v;
}

// Slab plan of the fused kernel: slabs of ~slab_kb KB, a whole number of unrolled trips, at most 64 per sample.
static void gnf_plan(int hw, int c, int esize, int rpb, int slab_kb, int* slabs, int* rows_per_slab) {
  const int trip = (gnf_ctas_per_sm() == 3 ? 4 : 8) * rpb;
  const long long row_bytes = (long long)c * esize;
  long long rows = ((long long)slab_kb * 1024 + row_bytes - 1) / row_bytes;
  rows = (rows + trip - 1) / trip * trip;
  if (rows > hw) rows = hw;
  if (rows < 1) rows = 1;
  int s = (int)((hw + rows - 1) / rows);
  while (s > 64) { rows += trip; s = (int)((hw + rows - 1) / rows); }
  *slabs = s;
  *rows_per_slab = (int)rows;
}

template <typename T>
static int gn_fused_launch(GnParams& P, int tpr, int rpb, int threads, cudaStream_t st) {
  const int ctas_per_sm = gnf_ctas_per_sm();
  static unsigned launch_no = 0;
  const int set = (int)(launch_no++ % kGnSyncSets);
  const size_t smem = (size_t)2 * rpb * P.c * sizeof(float);
  long long grid = (long long)P.n * P.slabs;
  const long long cap = (long long)ctas_per_sm * num_sms();
  if (grid > cap) grid = cap;
  if (ctas_per_sm == 3) gn_fused_kernel<T, 4, 3><<<(int)grid, threads, smem, st>>>(P, tpr, rpb, set);
  else gn_fused_kernel<T, 8, 2><<<(int)grid, threads, smem, st>>>(P, tpr, rpb, set);
  return check_cuda(cudaGetLastError(), "gn_fused_kernel launch");
}

// ---- one-pass GroupNorm: a thread-block cluster per sample, the sample resident in shared memory ------------
// The two-pass form reads x twice (statistics, apply): 3 HBM passes.  Here a cluster of 1..16 CTAs owns one
// sample: each CTA pulls its contiguous slab of rows into shared memory with 1-D bulk TMA copies (chunked, so
// the statistics run under the loads), the per-group partials of the CTAs meet through distributed shared
// memory in a fixed order (bit-reproducible), and the apply pass reads the slab from shared memory: 1 read +
// 1 write of HBM and one launch.  Used whenever hw*c*e / 16 fits a CTA's shared memory; the widest
// concatenated decoder inputs at 64x64 (5-8 MB per sample) stay on the two-pass kernels.
constexpr int kGnChunks = 8;

struct GnClusterSmem {
  uint64_t bar[kGnChunks];
  float red[2 * kGnMaxGroups];     // this CTA's per-group (sum, sumsq), read by the whole cluster
  float mean[kGnMaxGroups], rstd[kGnMaxGroups];
};

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               :: "r"(sm100::smem_u32(dst)), "l"(src), "r"(bytes), "r"(sm100::smem_u32(bar)) : "memory");
}

template <typename T, int CPT>
__global__ void __launch_bounds__(kGnThreads, 1)
gn_cluster_kernel(const GnParams P) {
  namespace cg = cooperative_groups;
  constexpr int E = V16<T>::E;
  extern __shared__ __align__(128) unsigned char gn_smem[];
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks();
  const int rank = (int)cluster.block_rank();
  const int n = blockIdx.x / cs;
  const int rows_per_cta = P.rows_per_slab;                 // ceil(hw / cs), set by the host
  const int r0 = min(P.hw, rank * rows_per_cta);
  const int nrows = min(P.hw, r0 + rows_per_cta) - r0;
  const int c1 = P.c1, c2 = P.c - P.c1;
  const GnMap M(P.c, E);
  // shared memory: header | slab 1 (rows x c1) | slab 2 (rows x c2) | per-(row lane, channel) partials
  GnClusterSmem* H = reinterpret_cast<GnClusterSmem*>(gn_smem);
  T* slab1 = reinterpret_cast<T*>(gn_smem + 1024);
  T* slab2 = slab1 + (size_t)rows_per_cta * c1;
  float* s_psum = reinterpret_cast<float*>(slab2 + (size_t)rows_per_cta * c2);
  float* s_psq = s_psum + M.rpb * P.c;

  const int rows_per_chunk = (nrows + kGnChunks - 1) / kGnChunks;
  if (threadIdx.x == 0) {
    for (int k = 0; k < kGnChunks; ++k) sm100::mbar_init(&H->bar[k], 1);
    sm100::fence_barrier_init();
    const T* g1 = reinterpret_cast<const T*>(P.x) + ((size_t)n * P.hw + r0) * c1;
    const T* g2 = P.x2 ? reinterpret_cast<const T*>(P.x2) + ((size_t)n * P.hw + r0) * c2 : nullptr;
    for (int k = 0; k < kGnChunks; ++k) {
      const int a = min(nrows, k * rows_per_chunk), b = min(nrows, (k + 1) * rows_per_chunk);
      const uint32_t b1 = (uint32_t)((size_t)(b - a) * c1 * sizeof(T));
      const uint32_t b2 = g2 ? (uint32_t)((size_t)(b - a) * c2 * sizeof(T)) : 0u;
      sm100::mbar_arrive_expect_tx(&H->bar[k], b1 + b2);      // 0 bytes: completes at once
      if (b1) bulk_g2s(slab1 + (size_t)a * c1, g1 + (size_t)a * c1, b1, &H->bar[k]);
      if (b2) bulk_g2s(slab2 + (size_t)a * c2, g2 + (size_t)a * c2, b2, &H->bar[k]);
    }
  }
  __syncthreads();

  const int cpg = P.c / P.groups;
  const T* add = P.add_nc ? reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c : nullptr;
  if (M.active) {
    float sum[CPT][E], sq[CPT][E], av[CPT][E];
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
#pragma unroll
      for (int j = 0; j < E; ++j) { sum[k][j] = 0.f; sq[k][j] = 0.f; av[k][j] = 0.f; }
      if (add && ch < P.c) V16<T>::ld(add + ch, av[k]);
    }
    for (int kc = 0; kc < kGnChunks; ++kc) {
      sm100::mbar_wait(&H->bar[kc], 0);
      const int a = min(nrows, kc * rows_per_chunk), b = min(nrows, (kc + 1) * rows_per_chunk);
      // rows a..b-1, strided over the row lanes so that every lane's row set does not depend on the chunking
      int r = a + ((M.my_row - a % M.rpb) + M.rpb) % M.rpb;
      for (; r < b; r += M.rpb) {
#pragma unroll
        for (int k = 0; k < CPT; ++k) {
          const int ch = (M.my_chunk + k * kGnThreads) * E;
          if (ch < P.c) {
            float v[E];
            V16<T>::ld(ch < c1 ? slab1 + (size_t)r * c1 + ch : slab2 + (size_t)r * c2 + (ch - c1), v);
#pragma unroll
            for (int j = 0; j < E; ++j) { const float t = v[j] + av[k][j]; sum[k][j] += t; sq[k][j] = fmaf(t, t, sq[k][j]); }
          }
        }
      }
    }
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
      if (ch < P.c) {
#pragma unroll
        for (int j = 0; j < E; ++j) {
          s_psum[M.my_row * P.c + ch + j] = sum[k][j];
          s_psq[M.my_row * P.c + ch + j] = sq[k][j];
        }
      }
    }
  } else {
    for (int kc = 0; kc < kGnChunks; ++kc) sm100::mbar_wait(&H->bar[kc], 0);
  }
  __syncthreads();
  if (threadIdx.x < P.groups) {
    float gs = 0.f, gq = 0.f;
    for (int r = 0; r < M.rpb; ++r)
      for (int ch = threadIdx.x * cpg; ch < (threadIdx.x + 1) * cpg; ++ch) {
        gs += s_psum[r * P.c + ch];
        gq += s_psq[r * P.c + ch];
      }
    H->red[2 * threadIdx.x] = gs;
    H->red[2 * threadIdx.x + 1] = gq;
  }
  cluster.sync();
  if (threadIdx.x < P.groups) {
    double s = 0.0, q = 0.0;
    for (int rk = 0; rk < cs; ++rk) {
      const float* remote = cluster.map_shared_rank(H->red, rk);
      s += (double)remote[2 * threadIdx.x];
      q += (double)remote[2 * threadIdx.x + 1];
    }
    const double cnt = (double)P.hw * cpg;
    const double mean = s / cnt;
    double var = q / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    H->mean[threadIdx.x] = (float)mean;
    H->rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)P.eps));
  }
  cluster.sync();       // all remote reads done (no CTA may exit before); mean/rstd visible to the CTA
  if (!M.active) return;
  T* y = reinterpret_cast<T*>(P.y) + ((size_t)n * P.hw + r0) * P.c;
  const T* gamma = reinterpret_cast<const T*>(P.gamma);
  const T* beta = reinterpret_cast<const T*>(P.beta);
  float scale[CPT][E], shift[CPT][E];
#pragma unroll
  for (int k = 0; k < CPT; ++k) {
    const int ch = (M.my_chunk + k * kGnThreads) * E;
    if (ch < P.c) {
      float g[E], b[E], a[E];
      V16<T>::ld(gamma + ch, g);
      V16<T>::ld(beta + ch, b);
#pragma unroll
      for (int j = 0; j < E; ++j) a[j] = 0.f;
      if (add) V16<T>::ld(add + ch, a);
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int grp = (ch + j) / cpg;
        scale[k][j] = H->rstd[grp] * g[j];
        shift[k][j] = fmaf(a[j] - H->mean[grp], scale[k][j], b[j]);
      }
    }
  }
  for (int r = M.my_row; r < nrows; r += M.rpb) {
#pragma unroll
    for (int k = 0; k < CPT; ++k) {
      const int ch = (M.my_chunk + k * kGnThreads) * E;
      if (ch < P.c) {
        float v[E];
        V16<T>::ld(ch < c1 ? slab1 + (size_t)r * c1 + ch : slab2 + (size_t)r * c2 + (ch - c1), v);
#pragma unroll
        for (int j = 0; j < E; ++j) {
          float t = fmaf(v[j], scale[k][j], shift[k][j]);
          if (P.silu) t = t / (1.0f + __expf(-t));
          v[j] = t;
        }
        V16<T>::st(y + (size_t)r * P.c + ch, v);
      }
    }
  }
}

// Cluster size for the one-pass kernel (0: does not fit, use the two-pass kernels).
static int gn_cluster_plan(int n, int hw, int c, int esize, int rpb, size_t* smem_bytes) {
  const size_t budget = 200 * 1024;
  for (int cs = 1; cs <= 16; cs *= 2) {
    const int rows = (hw + cs - 1) / cs;
    const size_t need = 1024 + (size_t)rows * c * esize + (size_t)2 * rpb * c * sizeof(float);
    if (need <= budget) {
      // enough CTAs to fill the chip if the sample can be cut further without making slabs tiny
      int best = cs;
      while (best < 16 && (long long)n * best < 2LL * num_sms() && (size_t)((hw + 2 * best - 1) / (2 * best)) * c * esize >= 16 * 1024) best *= 2;
      const int rows_b = (hw + best - 1) / best;
      *smem_bytes = 1024 + (size_t)rows_b * c * esize + (size_t)2 * rpb * c * sizeof(float);
      return best;
    }
  }
  return 0;
}

template <typename T, int CPT>
static int gn_cluster_launch(GnParams& P, int cs, size_t smem, cudaStream_t st) {
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gn_cluster_kernel<T, CPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024));
    VF_CUDA_TRY(cudaFuncSetAttribute(gn_cluster_kernel<T, CPT>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    attr = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(P.n * cs), 1, 1);
  cfg.blockDim = dim3(kGnThreads, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  // can a cluster of this size and footprint be co-scheduled at all?  (validated once per size class)
  static size_t ok_smem[17] = {0};
  static bool bad[17] = {false};
  if (bad[cs]) return -1;
  if (smem > ok_smem[cs]) {
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, gn_cluster_kernel<T, CPT>, &cfg) != cudaSuccess || n_clusters < 1) {
      cudaGetLastError();
      bad[cs] = true;
      return -1;                       // caller falls back to the two-pass kernels
    }
    ok_smem[cs] = smem;
  }
  return check_cuda(cudaLaunchKernelEx(&cfg, gn_cluster_kernel<T, CPT>, P), "gn_cluster_kernel launch");
}

// ---- resident GroupNorm: the fused persistent kernel with the slab held in shared memory ----------------------
// Same work items, tickets and per-sample counters as gn_fused_kernel, but a CTA pulls its slab (<= ~100 KB, a
// contiguous run of rows per source) into shared memory with 1-D bulk copies (cp.async.bulk, kGnrChunks pieces on
// their own mbarriers, so the statistics run under the loads and the whole slab is in flight at once, which no
// register-staged loop can afford), and the apply pass reads shared memory: x crosses HBM ONCE, y once.  Two CTAs
// per SM: while one streams its slab in, the other streams its result out.
constexpr int kGnrChunks = 8;
constexpr int kGnrMaxSlabs = 128;

constexpr int kGnrMaxThreads = kGnfMaxThreads;      // 512 threads per CTA (64 registers) measured slower: 0.146 vs 0.141 ms

struct GnrHeader {
  uint64_t bar[kGnrChunks];
  float mean[kGnMaxGroups], rstd[kGnMaxGroups];
  int item;
};
constexpr int kGnrHeaderBytes = 1024;
static_assert(sizeof(GnrHeader) <= kGnrHeaderBytes, "header");

template <typename T, int kMinB>
__global__ void __launch_bounds__(kGnfMaxThreads, kMinB)
gn_resident_kernel(const GnParams P, const int tpr, const int rpb, const int set) {
  constexpr int E = V16<T>::E;
  extern __shared__ __align__(128) unsigned char gnr_smem[];
  GnrHeader* H = reinterpret_cast<GnrHeader*>(gnr_smem);
  float4* s_pair = reinterpret_cast<float4*>(gnr_smem + kGnrHeaderBytes);          // per thread: (sum, sq) of its two groups
  T* slab1 = reinterpret_cast<T*>(gnr_smem + kGnrHeaderBytes + kGnrMaxThreads * sizeof(float4));
  unsigned int* sync = g_gn_sync[set & 0xff];
  const int tid = threadIdx.x;
  const int my_row = tid / tpr, my_chunk = tid - my_row * tpr;
  const bool active = my_row < rpb;
  const int ch = my_chunk * E;
  const int cpg = P.c / P.groups;                       // >= E: a 16-byte chunk touches at most two groups
  const int c1 = P.c1, c2 = P.c - P.c1;
  const int total = P.n * P.slabs;
  T* slab2 = slab1 + (size_t)P.rows_per_slab * c1;
  const bool from1 = ch < c1;
  const int ld = from1 ? c1 : c2;
  const T* my_slab = from1 ? slab1 + ch : slab2 + (ch - c1);
  const int g_lo = ch / cpg;
  const int n_lo = min(E, (g_lo + 1) * cpg - ch);       // channels of the chunk that belong to g_lo

  if (tid == 0) {
    for (int k = 0; k < kGnrChunks; ++k) sm100::mbar_init(&H->bar[k], 1);
    sm100::fence_barrier_init();
  }
  uint32_t phase = 0;
  // thread 0 takes the ticket of the NEXT item as soon as the current sample is complete (its round trip then hides
  // under the apply pass).  Not earlier: a ticket held while waiting could belong to the sample being waited for.
  int next_item = 0;
  if (tid == 0) next_item = (int)atomicAdd(&sync[0], 1u);
  for (;; phase ^= 1u) {
    if (tid == 0) H->item = next_item;
    __syncthreads();                                    // also: every thread is done with the previous slab
    const int item = H->item;
    if (item >= total) break;
    const int n = item / P.slabs, slab = item - n * P.slabs;
    const int r0 = slab * P.rows_per_slab;
    const int nrows = min(P.hw, r0 + P.rows_per_slab) - r0;
    int rows_per_chunk = (nrows + kGnrChunks - 1) / kGnrChunks;
    rows_per_chunk = (rows_per_chunk + rpb - 1) / rpb * rpb;        // a thread's rows do not depend on the chunking
    if (tid == 0) {
      const T* g1 = reinterpret_cast<const T*>(P.x) + ((size_t)n * P.hw + r0) * c1;
      const T* g2 = c2 ? reinterpret_cast<const T*>(P.x2) + ((size_t)n * P.hw + r0) * c2 : nullptr;
      for (int k = 0; k < kGnrChunks; ++k) {
        const int a = min(nrows, k * rows_per_chunk), b = min(nrows, (k + 1) * rows_per_chunk);
        const uint32_t b1 = (uint32_t)((size_t)(b - a) * c1 * sizeof(T));
        const uint32_t b2 = c2 ? (uint32_t)((size_t)(b - a) * c2 * sizeof(T)) : 0u;
        sm100::mbar_arrive_expect_tx(&H->bar[k], b1 + b2);            // 0 bytes: completes at once
        if (b1) bulk_g2s(slab1 + (size_t)a * c1, g1 + (size_t)a * c1, b1, &H->bar[k]);
        if (b2) bulk_g2s(slab2 + (size_t)a * c2, g2 + (size_t)a * c2, b2, &H->bar[k]);
      }
    }

    // parameters of phase 2 now: their latency hides under the slab load instead of sitting behind the sample wait
    float av[E];
#pragma unroll
    for (int j = 0; j < E; ++j) av[j] = 0.f;
    if (active && P.add_nc) V16<T>::ld(reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c + ch, av);

    // ---- phase 1: statistics as the chunks land (packed fp32 pairs: FADD2 / FFMA2) ---------------------
    {
      float sum[E], sq[E];
      {
        float2 sum2[E / 2], sq2[E / 2];
#pragma unroll
        for (int j = 0; j < E / 2; ++j) { sum2[j] = make_float2(0.f, 0.f); sq2[j] = make_float2(0.f, 0.f); }
        int r = my_row;
        for (int kc = 0; kc < kGnrChunks; ++kc) {
          sm100::mbar_wait(&H->bar[kc], phase);
          const int b = min(nrows, (kc + 1) * rows_per_chunk);
          if (active) {
#pragma unroll 4
            for (; r < b; r += rpb) {
              float2 v[E / 2];
              P16<T>::unpack(*reinterpret_cast<const uint4*>(my_slab + (size_t)r * ld), v);
#pragma unroll
              for (int j = 0; j < E / 2; ++j) { sum2[j] = __fadd2_rn(sum2[j], v[j]); sq2[j] = __ffma2_rn(v[j], v[j], sq2[j]); }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < E / 2; ++j) {
          sum[2 * j] = sum2[j].x; sum[2 * j + 1] = sum2[j].y;
          sq[2 * j] = sq2[j].x; sq[2 * j + 1] = sq2[j].y;
        }
      }
      if (active) {
        if (P.add_nc) {                                   // sum(x + a) = sum(x) + N a, sum((x + a)^2) = sum(x^2) + a (2 sum(x) + N a)
          const float cnt = (float)(my_row < nrows ? (nrows - my_row + rpb - 1) / rpb : 0);
#pragma unroll
          for (int j = 0; j < E; ++j) {
            sq[j] = fmaf(av[j], fmaf(cnt, av[j], 2.0f * sum[j]), sq[j]);
            sum[j] = fmaf(cnt, av[j], sum[j]);
          }
        }
        float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);      // (sum, sq) of g_lo, (sum, sq) of g_lo + 1
#pragma unroll
        for (int j = 0; j < E; ++j) {
          if (j < n_lo) { pr.x += sum[j]; pr.y += sq[j]; }
          else { pr.z += sum[j]; pr.w += sq[j]; }
        }
        s_pair[tid] = pr;
      }
    }
    __syncthreads();
    // group totals of the slab: eight lanes per group over (row lane, chunk) in a fixed order, then a fixed tree
    for (int grp = tid >> 3; grp < P.groups; grp += blockDim.x >> 3) {
      const int sub = tid & 7;
      const int k0 = (grp * cpg) / E, k1 = ((grp + 1) * cpg - 1) / E;       // chunks that touch the group
      const int nk = k1 - k0 + 1;
      float gs = 0.f, gq = 0.f;
      for (int i = sub; i < nk * rpb; i += 8) {
        const int rr = i / nk, k = k0 + (i - rr * nk);
        const float4 pr = s_pair[rr * tpr + k];
        const bool lo = (k * E) / cpg == grp;
        gs += lo ? pr.x : pr.z;
        gq += lo ? pr.y : pr.w;
      }
      const unsigned m8 = 0xffu << (tid & 24);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(m8, gs, o);
        gq += __shfl_xor_sync(m8, gq, o);
      }
      if (sub == 0) {
        float* w = P.ws + (((size_t)n * P.slabs + slab) * P.groups + grp) * 2;
        __stcg(w, gs);
        __stcg(w + 1, gq);
      }
    }
    __syncthreads();
    if (tid == 0) {
      // release: the CTA's partials (ordered before this thread by the barrier) become visible with the arrival
      unsigned seen;
      asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(seen) : "l"(&sync[2 + 2 * n]) : "memory");
      if (seen + 1u < (unsigned)P.slabs && !(set & 0x100))          // 0x100: VF_GN_DEBUG_NOWAIT (timing experiment, wrong results)
        while (ld_acquire_u32(&sync[2 + 2 * n]) < (unsigned)P.slabs) __nanosleep(32);
      else
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      next_item = (int)atomicAdd(&sync[0], 1u);
    }
    __syncthreads();
    // totals of the sample: eight lanes per group, slabs strided over the lanes (their loads go out together: one L2
    // round trip for up to 32 slabs), then a fixed shuffle tree in double -- the same order in every CTA of the sample
    if (tid < 8 * P.groups) {
      const int grp = tid >> 3, sub = tid & 7;
      double sd = 0.0, qd = 0.0;
      const float2* w = reinterpret_cast<const float2*>(P.ws) + ((size_t)n * P.slabs) * P.groups + grp;
      for (int i0 = sub; i0 < P.slabs; i0 += 32) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 8 * u;
          v[u] = i < P.slabs ? __ldcg(w + (size_t)i * P.groups) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { sd += (double)v[u].x; qd += (double)v[u].y; }
      }
      const unsigned m8 = 0xffu << (tid & 24);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sd += __shfl_xor_sync(m8, sd, o);
        qd += __shfl_xor_sync(m8, qd, o);
      }
      if (sub == 0) {
        const double cnt = (double)P.hw * cpg;
        const double mean = sd / cnt;
        double var = qd / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        H->mean[grp] = (float)mean;
        H->rstd[grp] = (float)(1.0 / sqrt(var + (double)P.eps));
      }
    }
    __syncthreads();

    // ---- phase 2: apply out of shared memory --------------------------------------------------------------
    if (active) {
      float scale[E], shift[E], gam[E], bet[E];          // gamma / beta: the same lines every item, L1 hits
      V16<T>::ld(reinterpret_cast<const T*>(P.gamma) + ch, gam);
      V16<T>::ld(reinterpret_cast<const T*>(P.beta) + ch, bet);
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int grp = (ch + j) / cpg;
        scale[j] = H->rstd[grp] * gam[j];
        shift[j] = fmaf(av[j] - H->mean[grp], scale[j], bet[j]);
      }
      // bf16 SiLU = h + h tanh(h), h = t / 2: the 1/2 is folded into scale / shift, the rest is two packed FMAs per pair
      constexpr bool kBf = sizeof(T) == 2;
      const float ks = (kBf && P.silu) ? 0.5f : 1.0f;
      float2 scale2[E / 2], shift2[E / 2];
#pragma unroll
      for (int j = 0; j < E / 2; ++j) {
        scale2[j] = make_float2(scale[2 * j] * ks, scale[2 * j + 1] * ks);
        shift2[j] = make_float2(shift[2 * j] * ks, shift[2 * j + 1] * ks);
      }
      char* q = reinterpret_cast<char*>(reinterpret_cast<T*>(P.y) + ((size_t)n * P.hw + r0 + my_row) * P.c + ch);
      const size_t qstride = (size_t)rpb * P.c * sizeof(T);
#pragma unroll 4
      for (int r = my_row; r < nrows; r += rpb, q += qstride) {
        float2 v[E / 2];
        P16<T>::unpack(*reinterpret_cast<const uint4*>(my_slab + (size_t)r * ld), v);
#pragma unroll
        for (int j = 0; j < E / 2; ++j) {
          float2 t = __ffma2_rn(v[j], scale2[j], shift2[j]);
          if (P.silu) {
            if constexpr (kBf) {
              float2 th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(t.x));
              asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(t.y));
              t = __ffma2_rn(t, th, t);
            } else {
              t = make_float2(silu_f<T>(t.x), silu_f<T>(t.y));
            }
          }
          v[j] = t;
        }
        st_na_v4(q, P16<T>::pack(v));
      }
    }
  }
  // the last CTA out re-arms the set: ticket, exit count and the arrival counters of all samples (every CTA has
  // left its loop by then, nobody reads them any more)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    H->item = (int)atomicAdd(&sync[1], 1u);
  }
  __syncthreads();
  if (H->item == (int)gridDim.x - 1) {
    for (int i = tid; i < P.n; i += blockDim.x) sync[2 + 2 * i] = 0u;
    if (tid == 0) { sync[0] = 0u; sync[1] = 0u; }
  }
}

// slab plan: the largest slab (a multiple of rpb rows) within `budget` bytes; 0 slabs = does not fit the scheme
static void gnr_plan(int n, int hw, int c, int esize, int rpb, size_t budget, int* slabs, int* rows_per_slab) {
  long long rows = (long long)(budget / ((size_t)c * esize));
  rows = rows / rpb * rpb;
  if (rows > hw) rows = (hw + rpb - 1) / rpb * rpb;
  if (rows < rpb) { *slabs = 0; *rows_per_slab = 0; return; }
  const int s_min = (int)((hw + rows - 1) / rows);
  // Default: the fewest slabs the budget allows.  VF_GN_RES_QUANT=1 searches a few more slab counts for the least
  // (rounds over 2 CTAs per SM) x (bytes per item + a synchronisation chain worth ~32 KB) -- with 96 samples of 7 slabs the
  // items make 2.27 rounds, three for the work of 2.27.  Measured equal or slightly slower (0.139 vs 0.132 ms at 96 x 4096 x
  // 320, profiles/r2_gn_ab.txt): tickets already balance the tail and the chain, not the rounds, bounds the kernel.
  static int quant = -1;
  if (quant < 0) { const char* e_ = getenv("VF_GN_RES_QUANT"); quant = e_ ? atoi(e_) : 0; }
  const long long ctas = 2LL * num_sms();
  int best_s = 0;
  long long best_rows = 0, best_cost = -1;
  for (int s = s_min; s <= (quant ? s_min + 12 : s_min) && s <= kGnrMaxSlabs; ++s) {
    long long r = ((hw + s - 1) / s + rpb - 1) / rpb * rpb;       // even slabs, a whole number of row lanes
    if (r < rpb || r * (long long)c * esize > (long long)budget) continue;
    const int s_eff = (int)((hw + r - 1) / r);
    const long long items = (long long)n * s_eff;
    const long long rounds = (items + ctas - 1) / ctas;
    // an item costs its bytes plus a synchronisation chain worth ~32 KB of streaming (three CTAs per SM with 64 KB slabs
    // measured 14 % slower than two with 100 KB); ties go to fewer slabs
    const long long cost = rounds * (r * (long long)c * esize + 32768) * 256 + s_eff;
    if (best_cost < 0 || cost < best_cost) { best_cost = cost; best_s = s_eff; best_rows = r; }
  }
  *slabs = best_s;
  *rows_per_slab = (int)best_rows;
}

// Two CTAs per SM, ~100 KB slabs (three with ~64 KB slabs measured slower: 0.152 vs 0.135 ms at 96 x 4096 x 320 -- the
// per-item synchronisation is amortised over fewer bytes).
static size_t gnr_smem_per_cta() { return 112 * 1024; }

template <typename T, int kMinB>
static int gn_resident_launch_t(GnParams& P, int tpr, int rpb, int threads, cudaStream_t st) {
  static unsigned launch_no = 0;
  const int set = (int)(launch_no++ % kGnSyncSets);
  const size_t smem = kGnrHeaderBytes + kGnrMaxThreads * sizeof(float4) + (size_t)P.rows_per_slab * P.c * sizeof(T);
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gn_resident_kernel<T, kMinB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnr_smem_per_cta()));
    attr = true;
  }
  long long grid = (long long)P.n * P.slabs;
  const long long cap = (long long)kMinB * num_sms();
  if (grid > cap) grid = cap;
  static int nowait = -1;
  if (nowait < 0) { const char* e_ = getenv("VF_GN_DEBUG_NOWAIT"); nowait = e_ ? atoi(e_) : 0; }
  gn_resident_kernel<T, kMinB><<<(int)grid, threads, smem, st>>>(P, tpr, rpb, set | (nowait ? 0x100 : 0));
  return check_cuda(cudaGetLastError(), "gn_resident_kernel launch");
}
template <typename T>
static int gn_resident_launch(GnParams& P, int tpr, int rpb, int threads, cudaStream_t st) {
  return gn_resident_launch_t<T, 2>(P, tpr, rpb, threads, st);
}

// ---- resident GroupNorm, two slabs per CTA in flight (third session of round 2) ---------------------------------
// gn_resident_kernel runs the phases of an item strictly in sequence -- load, statistics, publish, WAIT for the rest of the
// sample, totals, apply -- and its CTA moves no data while it waits: 23 % of all warp-stall samples sit on the barrier behind
// thread 0's poll of the sample counter, another 10 % on the slab's own load (ncu source view of profiles/r2_gn_resident_ncu.txt).
// Here a CTA holds TWO half-size slabs: the FRONT half of an item (issue the bulk copies, statistics as they land, publish the
// partials, arrive on the sample counter) never blocks on another CTA, so the front of item i+1 is run BEFORE the BACK half of
// item i (wait for the sample, totals, apply out of shared memory): the wait of item i and the load of item i+1 overlap, and
// the copies of item i+2 are in flight under the apply pass of item i+1.  MEASURED SLOWER than the one-slab kernel (see the
// dispatch below); kept as an opt-in (VF_GN_PIPE=1).  Deadlock-free for the same reason as before: every
// ticket a CTA holds is published before that CTA waits for anything (tickets are handed out in sample-major order, so every
// slab of an awaited sample is held by a running CTA that will publish it without waiting).
struct Gnr2Header {
  uint64_t bar[2][kGnrChunks];
  float mean[kGnMaxGroups], rstd[kGnMaxGroups];
  int item;
};
static_assert(sizeof(Gnr2Header) <= kGnrHeaderBytes, "header");

template <typename T, int kMinB>
__global__ void __launch_bounds__(kGnfMaxThreads, kMinB)
gn_resident2_kernel(const GnParams P, const int tpr, const int rpb, const int set) {
  constexpr int E = V16<T>::E;
  extern __shared__ __align__(128) unsigned char gnr_smem[];
  Gnr2Header* H = reinterpret_cast<Gnr2Header*>(gnr_smem);
  float4* s_pair = reinterpret_cast<float4*>(gnr_smem + kGnrHeaderBytes);          // per thread: (sum, sq) of its two groups
  T* buf0 = reinterpret_cast<T*>(gnr_smem + kGnrHeaderBytes + kGnrMaxThreads * sizeof(float4));
  unsigned int* sync = g_gn_sync[set & 0xff];
  const int tid = threadIdx.x;
  const int my_row = tid / tpr, my_chunk = tid - my_row * tpr;
  const bool active = my_row < rpb;
  const int ch = my_chunk * E;
  const int cpg = P.c / P.groups;                       // >= E: a 16-byte chunk touches at most two groups
  const int c1 = P.c1, c2 = P.c - P.c1;
  const int total = P.n * P.slabs;
  const size_t buf_elems = (size_t)P.rows_per_slab * P.c;
  const bool from1 = ch < c1;
  const int ld = from1 ? c1 : c2;
  const size_t my_off = from1 ? (size_t)ch : (size_t)P.rows_per_slab * c1 + (ch - c1);     // within a buffer: slab1 | slab2
  const int g_lo = ch / cpg;
  const int n_lo = min(E, (g_lo + 1) * cpg - ch);       // channels of the chunk that belong to g_lo

  if (tid == 0) {
    for (int b = 0; b < 2; ++b)
      for (int k = 0; k < kGnrChunks; ++k) sm100::mbar_init(&H->bar[b][k], 1);
    sm100::fence_barrier_init();
  }

  // next ticket, broadcast to the CTA.  The barrier also orders every thread's reads of a buffer (apply pass) before
  // thread 0 hands that buffer to the next bulk copies.
  auto take = [&]() -> int {
    if (tid == 0) H->item = (int)atomicAdd(&sync[0], 1u);
    __syncthreads();
    const int it = H->item;
    __syncthreads();                                    // H->item may be rewritten by the next take()
    return it;
  };

  // ---- front half: loads, statistics, publish, arrive.  Never waits for another CTA. ---------------------------------
  auto front = [&](const int item, const int b, const uint32_t phase, float (&av)[E]) {
    const int n = item / P.slabs, slab = item - n * P.slabs;
    const int r0 = slab * P.rows_per_slab;
    const int nrows = min(P.hw, r0 + P.rows_per_slab) - r0;
    int rows_per_chunk = (nrows + kGnrChunks - 1) / kGnrChunks;
    rows_per_chunk = (rows_per_chunk + rpb - 1) / rpb * rpb;        // a thread's rows do not depend on the chunking
    T* slab1 = buf0 + (size_t)b * buf_elems;
    T* slab2 = slab1 + (size_t)P.rows_per_slab * c1;
    if (tid == 0) {
      const T* g1 = reinterpret_cast<const T*>(P.x) + ((size_t)n * P.hw + r0) * c1;
      const T* g2 = c2 ? reinterpret_cast<const T*>(P.x2) + ((size_t)n * P.hw + r0) * c2 : nullptr;
      for (int k = 0; k < kGnrChunks; ++k) {
        const int a = min(nrows, k * rows_per_chunk), e = min(nrows, (k + 1) * rows_per_chunk);
        const uint32_t b1 = (uint32_t)((size_t)(e - a) * c1 * sizeof(T));
        const uint32_t b2 = c2 ? (uint32_t)((size_t)(e - a) * c2 * sizeof(T)) : 0u;
        sm100::mbar_arrive_expect_tx(&H->bar[b][k], b1 + b2);          // 0 bytes: completes at once
        if (b1) bulk_g2s(slab1 + (size_t)a * c1, g1 + (size_t)a * c1, b1, &H->bar[b][k]);
        if (b2) bulk_g2s(slab2 + (size_t)a * c2, g2 + (size_t)a * c2, b2, &H->bar[b][k]);
      }
    }
#pragma unroll
    for (int j = 0; j < E; ++j) av[j] = 0.f;
    if (active && P.add_nc) V16<T>::ld(reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c + ch, av);
    const T* my_slab = slab1 + my_off;
    {
      float sum[E], sq[E];
      {
        float2 sum2[E / 2], sq2[E / 2];
#pragma unroll
        for (int j = 0; j < E / 2; ++j) { sum2[j] = make_float2(0.f, 0.f); sq2[j] = make_float2(0.f, 0.f); }
        int r = my_row;
        for (int kc = 0; kc < kGnrChunks; ++kc) {
          sm100::mbar_wait(&H->bar[b][kc], phase);
          const int e = min(nrows, (kc + 1) * rows_per_chunk);
          if (active) {
#pragma unroll 4
            for (; r < e; r += rpb) {
              float2 v[E / 2];
              P16<T>::unpack(*reinterpret_cast<const uint4*>(my_slab + (size_t)r * ld), v);
#pragma unroll
              for (int j = 0; j < E / 2; ++j) { sum2[j] = __fadd2_rn(sum2[j], v[j]); sq2[j] = __ffma2_rn(v[j], v[j], sq2[j]); }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < E / 2; ++j) {
          sum[2 * j] = sum2[j].x; sum[2 * j + 1] = sum2[j].y;
          sq[2 * j] = sq2[j].x; sq[2 * j + 1] = sq2[j].y;
        }
      }
      if (active) {
        if (P.add_nc) {                                   // sum(x + a) = sum(x) + N a, sum((x + a)^2) = sum(x^2) + a (2 sum(x) + N a)
          const float cnt = (float)(my_row < nrows ? (nrows - my_row + rpb - 1) / rpb : 0);
#pragma unroll
          for (int j = 0; j < E; ++j) {
            sq[j] = fmaf(av[j], fmaf(cnt, av[j], 2.0f * sum[j]), sq[j]);
            sum[j] = fmaf(cnt, av[j], sum[j]);
          }
        }
        float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);      // (sum, sq) of g_lo, (sum, sq) of g_lo + 1
#pragma unroll
        for (int j = 0; j < E; ++j) {
          if (j < n_lo) { pr.x += sum[j]; pr.y += sq[j]; }
          else { pr.z += sum[j]; pr.w += sq[j]; }
        }
        s_pair[tid] = pr;
      }
    }
    __syncthreads();
    // group totals of the slab: eight lanes per group over (row lane, chunk) in a fixed order, then a fixed tree
    for (int grp = tid >> 3; grp < P.groups; grp += blockDim.x >> 3) {
      const int sub = tid & 7;
      const int k0 = (grp * cpg) / E, k1 = ((grp + 1) * cpg - 1) / E;       // chunks that touch the group
      const int nk = k1 - k0 + 1;
      float gs = 0.f, gq = 0.f;
      for (int i = sub; i < nk * rpb; i += 8) {
        const int rr = i / nk, k = k0 + (i - rr * nk);
        const float4 pr = s_pair[rr * tpr + k];
        const bool lo = (k * E) / cpg == grp;
        gs += lo ? pr.x : pr.z;
        gq += lo ? pr.y : pr.w;
      }
      const unsigned m8 = 0xffu << (tid & 24);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        gs += __shfl_xor_sync(m8, gs, o);
        gq += __shfl_xor_sync(m8, gq, o);
      }
      if (sub == 0) {
        float* w = P.ws + (((size_t)n * P.slabs + slab) * P.groups + grp) * 2;
        __stcg(w, gs);
        __stcg(w + 1, gq);
      }
    }
    __syncthreads();
    if (tid == 0) {
      // release: the CTA's partials (ordered before this thread by the barrier) become visible with the arrival
      unsigned seen;
      asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(seen) : "l"(&sync[2 + 2 * n]) : "memory");
      (void)seen;
    }
  };

  // ---- back half: wait for the sample, totals, apply out of shared memory ----------------------------------------------
  auto back = [&](const int item, const int b, const float (&av)[E]) {
    const int n = item / P.slabs, slab = item - n * P.slabs;
    const int r0 = slab * P.rows_per_slab;
    const int nrows = min(P.hw, r0 + P.rows_per_slab) - r0;
    const T* my_slab = buf0 + (size_t)b * buf_elems + my_off;
    if (tid == 0) {
      if (!(set & 0x100))                                 // 0x100: VF_GN_DEBUG_NOWAIT (timing experiment, wrong results)
        while (ld_acquire_u32(&sync[2 + 2 * n]) < (unsigned)P.slabs) __nanosleep(32);
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
    }
    __syncthreads();
    // totals of the sample: eight lanes per group, slabs strided over the lanes, then a fixed shuffle tree in double -- the
    // same order in every CTA of the sample
    if (tid < 8 * P.groups) {
      const int grp = tid >> 3, sub = tid & 7;
      double sd = 0.0, qd = 0.0;
      const float2* w = reinterpret_cast<const float2*>(P.ws) + ((size_t)n * P.slabs) * P.groups + grp;
      for (int i0 = sub; i0 < P.slabs; i0 += 32) {
        float2 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int i = i0 + 8 * u;
          v[u] = i < P.slabs ? __ldcg(w + (size_t)i * P.groups) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) { sd += (double)v[u].x; qd += (double)v[u].y; }
      }
      const unsigned m8 = 0xffu << (tid & 24);
#pragma unroll
      for (int o = 4; o > 0; o >>= 1) {
        sd += __shfl_xor_sync(m8, sd, o);
        qd += __shfl_xor_sync(m8, qd, o);
      }
      if (sub == 0) {
        const double cnt = (double)P.hw * cpg;
        const double mean = sd / cnt;
        double var = qd / cnt - mean * mean;
        if (var < 0.0) var = 0.0;
        H->mean[grp] = (float)mean;
        H->rstd[grp] = (float)(1.0 / sqrt(var + (double)P.eps));
      }
    }
    __syncthreads();
    if (active) {
      float scale[E], shift[E], gam[E], bet[E];          // gamma / beta: the same lines every item, L1 hits
      V16<T>::ld(reinterpret_cast<const T*>(P.gamma) + ch, gam);
      V16<T>::ld(reinterpret_cast<const T*>(P.beta) + ch, bet);
#pragma unroll
      for (int j = 0; j < E; ++j) {
        const int grp = (ch + j) / cpg;
        scale[j] = H->rstd[grp] * gam[j];
        shift[j] = fmaf(av[j] - H->mean[grp], scale[j], bet[j]);
      }
      // bf16 SiLU = h + h tanh(h), h = t / 2: the 1/2 is folded into scale / shift, the rest is two packed FMAs per pair
      constexpr bool kBf = sizeof(T) == 2;
      const float ks = (kBf && P.silu) ? 0.5f : 1.0f;
      float2 scale2[E / 2], shift2[E / 2];
#pragma unroll
      for (int j = 0; j < E / 2; ++j) {
        scale2[j] = make_float2(scale[2 * j] * ks, scale[2 * j + 1] * ks);
        shift2[j] = make_float2(shift[2 * j] * ks, shift[2 * j + 1] * ks);
      }
      char* q = reinterpret_cast<char*>(reinterpret_cast<T*>(P.y) + ((size_t)n * P.hw + r0 + my_row) * P.c + ch);
      const size_t qstride = (size_t)rpb * P.c * sizeof(T);
#pragma unroll 4
      for (int r = my_row; r < nrows; r += rpb, q += qstride) {
        float2 v[E / 2];
        P16<T>::unpack(*reinterpret_cast<const uint4*>(my_slab + (size_t)r * ld), v);
#pragma unroll
        for (int j = 0; j < E / 2; ++j) {
          float2 t = __ffma2_rn(v[j], scale2[j], shift2[j]);
          if (P.silu) {
            if constexpr (kBf) {
              float2 th;
              asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(t.x));
              asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(t.y));
              t = __ffma2_rn(t, th, t);
            } else {
              t = make_float2(silu_f<T>(t.x), silu_f<T>(t.y));
            }
          }
          v[j] = t;
        }
        st_na_v4(q, P16<T>::pack(v));
      }
    }
  };

  float av_a[E], av_b[E];
  uint32_t ph_a = 0, ph_b = 0;
  __syncthreads();                                      // barrier initialisation visible
  int it_a = take();
  if (it_a < total) {
    front(it_a, 0, ph_a, av_a); ph_a ^= 1u;
    for (;;) {
      const int it_b = take();
      if (it_b < total) { front(it_b, 1, ph_b, av_b); ph_b ^= 1u; }
      back(it_a, 0, av_a);
      if (it_b >= total) break;
      it_a = take();
      if (it_a < total) { front(it_a, 0, ph_a, av_a); ph_a ^= 1u; }
      back(it_b, 1, av_b);
      if (it_a >= total) break;
    }
  }
  // the last CTA out re-arms the set: ticket, exit count and the arrival counters of all samples (every CTA has
  // left its loop by then, nobody reads them any more)
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    H->item = (int)atomicAdd(&sync[1], 1u);
  }
  __syncthreads();
  if (H->item == (int)gridDim.x - 1) {
    for (int i = tid; i < P.n; i += blockDim.x) sync[2 + 2 * i] = 0u;
    if (tid == 0) { sync[0] = 0u; sync[1] = 0u; }
  }
}

template <typename T>
static int gn_resident2_launch(GnParams& P, int tpr, int rpb, int threads, cudaStream_t st) {
  static unsigned launch_no2 = 0;
  const int set = (int)(launch_no2++ % kGnSyncSets);
  const size_t smem = kGnrHeaderBytes + kGnrMaxThreads * sizeof(float4) + 2 * (size_t)P.rows_per_slab * P.c * sizeof(T);
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gn_resident2_kernel<T, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gnr_smem_per_cta()));
    attr = true;
  }
  long long grid = (long long)P.n * P.slabs;
  const long long cap = 2LL * num_sms();
  if (grid > cap) grid = cap;
  static int nowait = -1;
  if (nowait < 0) { const char* e_ = getenv("VF_GN_DEBUG_NOWAIT"); nowait = e_ ? atoi(e_) : 0; }
  gn_resident2_kernel<T, 2><<<(int)grid, threads, smem, st>>>(P, tpr, rpb, set | (nowait ? 0x100 : 0));
  return check_cuda(cudaGetLastError(), "gn_resident2_kernel launch");
}

// ---- resident GroupNorm, warp-specialised (bf16; third session of round 2) ----------------------------------------
// The one-slab kernel runs an item's phases in sequence on ALL its warps, with seven block barriers and a chain that ONE thread
// walks (atomic arrival, poll, fence, ticket) while the others park; two slabs per CTA did not help because that fixed cost is
// paid per item (gn_resident2_kernel).  Here the phases are ROLES, each with its own warps, meeting only on mbarriers:
//   warp 0        producer: waits for a free buffer, THEN takes a ticket (a held ticket is always loadable at once), issues the
//                 bulk copies of the slab into a ring of three 66 KB buffers;
//   warps 1-8     statistics: per-thread sums out of shared memory, slab totals in a fixed order, partials to L2, release-arrive
//                 on the sample counter.  This group never waits for another CTA.  (A first version let the LAST arrival of a
//                 sample reduce the partials once and raise a ready flag: arrival -> reduce -> flag -> poll -> read is five L2
//                 round trips on every item's critical path, 0.177 ms against 0.125 with the waits skipped.)
//   warps 9-24    apply, two groups of eight warps taking alternate items (one group is latency-bound: 2.7 us per 66 KB slab
//                 against 3 us of HBM time): wait until the sample's arrival counter is full (one thread polls), sum its partials
//                 (eight lanes per group, one L2 round trip), normalise (+ SiLU) out of shared memory, store, hand the buffer back.
// One CTA per SM; the copies of item i+2 are in flight while item i+1 is summed and item i is applied.  Deadlock-free: a ticket
// is taken only when its buffer is free, so every held ticket is published without waiting on anyone; 3 x grid tickets in
// flight always cover the oldest awaited sample (<= 128 slabs).  Bit-reproducible: fixed orders everywhere, and the totals of a
// sample are computed exactly once.
constexpr int kGwsBufs = 3;
constexpr int kGwsGroup = 256;                          // threads of the statistics group and of the apply group
constexpr int kGwsApplyGroups = 2;                      // apply groups taking alternate items (the apply pass is latency-bound at 8 warps)
constexpr int kGwsThreads = 32 + (1 + kGwsApplyGroups) * kGwsGroup;
constexpr size_t kGwsBufBytes = 66 * 1024;
constexpr int kGwsSlabStride = 128;                     // workspace: [n][<= 128 slabs][groups][2] partials, then [n][groups][2] totals

struct GwsHeader {
  uint64_t full[kGwsBufs], stats_done[kGwsBufs], empty[kGwsBufs];
  int item[kGwsBufs];
  float2 mr[2][kGnMaxGroups];                           // per apply group: (mean, rstd) of the sample being applied
};
static_assert(sizeof(GwsHeader) <= 1024, "header");

__device__ __forceinline__ void gws_bar(int id) { asm volatile("bar.sync %0, %1;" :: "r"(id), "n"(kGwsGroup) : "memory"); }

__global__ void __launch_bounds__(kGwsThreads, 1)
gn_ws_kernel(const GnParams P, const int tpr, const int rpb, const int set) {
  using T = __nv_bfloat16;
  constexpr int E = 8;
  extern __shared__ __align__(128) unsigned char gws_smem[];
  GwsHeader* H = reinterpret_cast<GwsHeader*>(gws_smem);
  float4* s_pair = reinterpret_cast<float4*>(gws_smem + 1024);
  unsigned char* bufs = gws_smem + 1024 + kGwsGroup * sizeof(float4);
  unsigned int* sync = g_gn_sync[set & 0xff];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c1 = P.c1, c2 = P.c - P.c1;
  const int cpg = P.c / P.groups;
  const int total = P.n * P.slabs;

  if (threadIdx.x == 0) {
    for (int b = 0; b < kGwsBufs; ++b) {
      sm100::mbar_init(&H->full[b], 1);
      sm100::mbar_init(&H->stats_done[b], 1);
      sm100::mbar_init(&H->empty[b], 1);
    }
    sm100::fence_barrier_init();
  }
  __syncthreads();

  if (warp == 0) {
    // =========================== producer ================================================================
    if (lane == 0) {
      int ends = 0;
      for (int i = 0;; ++i) {
        const int b = i % kGwsBufs;
        sm100::mbar_wait(&H->empty[b], ((uint32_t)(i / kGwsBufs) & 1u) ^ 1u);      // buffer free FIRST, ticket second
        const int item = ends ? total : (int)atomicAdd(&sync[0], 1u);
        H->item[b] = item;
        if (item >= total) {                                                       // end markers, no data: one per apply group
          sm100::mbar_arrive(&H->full[b]);
          if (++ends == kGwsApplyGroups) break;
          continue;
        }
        const int n = item / P.slabs, slab = item - n * P.slabs;
        const int r0 = slab * P.rows_per_slab;
        const int nrows = min(P.hw, r0 + P.rows_per_slab) - r0;
        T* slab1 = reinterpret_cast<T*>(bufs + (size_t)b * kGwsBufBytes);
        T* slab2 = slab1 + (size_t)P.rows_per_slab * c1;
        const T* g1 = reinterpret_cast<const T*>(P.x) + ((size_t)n * P.hw + r0) * c1;
        const uint32_t b1 = (uint32_t)((size_t)nrows * c1 * sizeof(T));
        const uint32_t b2 = c2 ? (uint32_t)((size_t)nrows * c2 * sizeof(T)) : 0u;
        sm100::mbar_arrive_expect_tx(&H->full[b], b1 + b2);
        // pieces of <= 16 KB: several copies in flight per slab
        for (uint32_t o = 0; o < b1; o += 16384u)
          bulk_g2s(reinterpret_cast<char*>(slab1) + o, reinterpret_cast<const char*>(g1) + o, min(16384u, b1 - o), &H->full[b]);
        if (b2) {
          const T* g2 = reinterpret_cast<const T*>(P.x2) + ((size_t)n * P.hw + r0) * c2;
          for (uint32_t o = 0; o < b2; o += 16384u)
            bulk_g2s(reinterpret_cast<char*>(slab2) + o, reinterpret_cast<const char*>(g2) + o, min(16384u, b2 - o), &H->full[b]);
        }
      }
    }
  } else {
    const bool is_stats = warp <= 8;
    const int ag = is_stats ? 0 : (warp - 9) >> 3;                           // apply group
    const int gt = threadIdx.x - 32 - (is_stats ? 0 : (1 + ag) * kGwsGroup); // 0..255 within the group
    const int my_row = gt / tpr, my_chunk = gt - my_row * tpr;
    const bool active = my_row < rpb;
    const int ch = my_chunk * E;
    const bool from1 = ch < c1;
    const int ld = from1 ? c1 : c2;
    const size_t my_off = from1 ? (size_t)ch : (size_t)P.rows_per_slab * c1 + (ch - c1);
    const int g_lo = ch / cpg;
    const int n_lo = min(E, (g_lo + 1) * cpg - ch);       // channels of the chunk that belong to g_lo

    if (is_stats) {
      // =========================== statistics ============================================================
      int ends = 0;
      for (int i = 0;; ++i) {
        const int b = i % kGwsBufs;
        sm100::mbar_wait(&H->full[b], (uint32_t)(i / kGwsBufs) & 1u);
        const int item = H->item[b];
        if (item >= total) {
          if (gt == 0) sm100::mbar_arrive(&H->stats_done[b]);
          if (++ends == kGwsApplyGroups) break;
          continue;
        }
        const int n = item / P.slabs, slab = item - n * P.slabs;
        const int r0 = slab * P.rows_per_slab;
        const int nrows = min(P.hw, r0 + P.rows_per_slab) - r0;
        const T* my_slab = reinterpret_cast<const T*>(bufs + (size_t)b * kGwsBufBytes) + my_off;
        float av[E];
#pragma unroll
        for (int j = 0; j < E; ++j) av[j] = 0.f;
        if (active && P.add_nc) V16<T>::ld(reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c + ch, av);
        if (active) {
          float2 sum2[E / 2], sq2[E / 2];
#pragma unroll
          for (int j = 0; j < E / 2; ++j) { sum2[j] = make_float2(0.f, 0.f); sq2[j] = make_float2(0.f, 0.f); }
#pragma unroll 4
          for (int r = my_row; r < nrows; r += rpb) {
            float2 v[E / 2];
            P16<T>::unpack(*reinterpret_cast<const uint4*>(my_slab + (size_t)r * ld), v);
#pragma unroll
            for (int j = 0; j < E / 2; ++j) { sum2[j] = __fadd2_rn(sum2[j], v[j]); sq2[j] = __ffma2_rn(v[j], v[j], sq2[j]); }
          }
          float sum[E], sq[E];
#pragma unroll
          for (int j = 0; j < E / 2; ++j) {
            sum[2 * j] = sum2[j].x; sum[2 * j + 1] = sum2[j].y;
            sq[2 * j] = sq2[j].x; sq[2 * j + 1] = sq2[j].y;
          }
          if (P.add_nc) {                                 // sum(x + a) = sum(x) + N a, sum((x + a)^2) = sum(x^2) + a (2 sum(x) + N a)
            const float cnt = (float)(my_row < nrows ? (nrows - my_row + rpb - 1) / rpb : 0);
#pragma unroll
            for (int j = 0; j < E; ++j) {
              sq[j] = fmaf(av[j], fmaf(cnt, av[j], 2.0f * sum[j]), sq[j]);
              sum[j] = fmaf(cnt, av[j], sum[j]);
            }
          }
          float4 pr = make_float4(0.f, 0.f, 0.f, 0.f);    // (sum, sq) of g_lo, (sum, sq) of g_lo + 1
#pragma unroll
          for (int j = 0; j < E; ++j) {
            if (j < n_lo) { pr.x += sum[j]; pr.y += sq[j]; }
            else { pr.z += sum[j]; pr.w += sq[j]; }
          }
          s_pair[gt] = pr;
        }
        gws_bar(1);
        // slab totals per group: eight lanes per group over (row lane, chunk) in a fixed order, then a fixed tree
        {
          const int grp = gt >> 3, sub = gt & 7;
          if (grp < P.groups) {
            const int k0 = (grp * cpg) / E, k1 = ((grp + 1) * cpg - 1) / E;       // chunks that touch the group
            const int nk = k1 - k0 + 1;
            float gs = 0.f, gq = 0.f;
            for (int q = sub; q < nk * rpb; q += 8) {
              const int rr = q / nk, k = k0 + (q - rr * nk);
              const float4 pr = s_pair[rr * tpr + k];
              const bool lo = (k * E) / cpg == grp;
              gs += lo ? pr.x : pr.z;
              gq += lo ? pr.y : pr.w;
            }
            const unsigned m8 = 0xffu << (lane & 24);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
              gs += __shfl_xor_sync(m8, gs, o);
              gq += __shfl_xor_sync(m8, gq, o);
            }
            if (sub == 0) __stcg(reinterpret_cast<float2*>(P.ws) + ((size_t)n * P.slabs + slab) * P.groups + grp, make_float2(gs, gq));
          }
        }
        gws_bar(1);                                       // every partial of this slab written; s_pair free again
        if (gt == 0) {
          // release: the slab's partials (ordered before this thread by the barrier) become visible with the arrival
          unsigned seen;
          asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(seen) : "l"(&sync[2 + 2 * n]) : "memory");
          (void)seen;
          sm100::mbar_arrive(&H->stats_done[b]);
        }
      }
    } else {
      // =========================== apply ====================================================================
      float gam[E], bet[E];                               // the same for every item
#pragma unroll
      for (int j = 0; j < E; ++j) { gam[j] = 0.f; bet[j] = 0.f; }
      if (active) {
        V16<T>::ld(reinterpret_cast<const T*>(P.gamma) + ch, gam);
        V16<T>::ld(reinterpret_cast<const T*>(P.beta) + ch, bet);
      }
      const bool bulk_out = (set & 0x200) && c2 == 0;
      for (int i = ag;; i += kGwsApplyGroups) {
        const int b = i % kGwsBufs;
        sm100::mbar_wait(&H->stats_done[b], (uint32_t)(i / kGwsBufs) & 1u);
        const int item = H->item[b];
        if (item >= total) break;
        const int n = item / P.slabs, slab = item - n * P.slabs;
        const int r0 = slab * P.rows_per_slab;
        const int nrows = min(P.hw, r0 + P.rows_per_slab) - r0;
        const T* my_slab = reinterpret_cast<const T*>(bufs + (size_t)b * kGwsBufBytes) + my_off;
        float av[E];                                        // issued before the wait: needs only the sample index
#pragma unroll
        for (int j = 0; j < E; ++j) av[j] = 0.f;
        if (active && P.add_nc) V16<T>::ld(reinterpret_cast<const T*>(P.add_nc) + (size_t)n * P.c + ch, av);
        if (gt == 0) {
          if (!(set & 0x100))                               // 0x100: VF_GN_DEBUG_NOWAIT (timing experiment, wrong results)
            while (ld_acquire_u32(&sync[2 + 2 * n]) < (unsigned)P.slabs) __nanosleep(32);
          asm volatile("fence.acq_rel.gpu;" ::: "memory");
        }
        gws_bar(2 + ag);
        // totals of the sample: eight lanes per group, a lane's slabs (s = sub, sub + 8, ...) all in flight at once (one L2 round
        // trip), summed in that order, then a fixed tree in double -- the same order in every CTA of the sample
        {
          const int grp = gt >> 3, sub = gt & 7;
          if (grp < P.groups) {
            const float2* w = reinterpret_cast<const float2*>(P.ws) + ((size_t)n * P.slabs) * P.groups + grp;
            float2 v[kGwsSlabStride / 8];
#pragma unroll
            for (int u = 0; u < kGwsSlabStride / 8; ++u) {
              const int sl = sub + 8 * u;
              v[u] = sl < P.slabs ? __ldcg(w + (size_t)sl * P.groups) : make_float2(0.f, 0.f);
            }
            double sd = 0.0, qd = 0.0;
#pragma unroll
            for (int u = 0; u < kGwsSlabStride / 8; ++u) { sd += (double)v[u].x; qd += (double)v[u].y; }
            const unsigned m8 = 0xffu << (lane & 24);
#pragma unroll
            for (int o = 4; o > 0; o >>= 1) {
              sd += __shfl_xor_sync(m8, sd, o);
              qd += __shfl_xor_sync(m8, qd, o);
            }
            if (sub == 0) {
              const double cnt = (double)P.hw * cpg;
              const double mean = sd / cnt;
              double var = qd / cnt - mean * mean;
              if (var < 0.0) var = 0.0;
              H->mr[ag][grp] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)P.eps)));
            }
          }
        }
        gws_bar(2 + ag);
        if (active) {
          const float2 t_lo = H->mr[ag][g_lo];
          const float2 t_hi = n_lo < E ? H->mr[ag][g_lo + 1] : t_lo;
          // bf16 SiLU = h + h tanh(h), h = t / 2: the 1/2 is folded into scale / shift, the rest is two packed FMAs per pair
          const float ks = P.silu ? 0.5f : 1.0f;
          float scale[E], shift[E];
#pragma unroll
          for (int j = 0; j < E; ++j) {
            const float2 t = j < n_lo ? t_lo : t_hi;
            scale[j] = t.y * gam[j];
            shift[j] = fmaf(av[j] - t.x, scale[j], bet[j]) * ks;
            scale[j] *= ks;
          }
          float2 scale2[E / 2], shift2[E / 2];
#pragma unroll
          for (int j = 0; j < E / 2; ++j) {
            scale2[j] = make_float2(scale[2 * j], scale[2 * j + 1]);
            shift2[j] = make_float2(shift[2 * j], shift[2 * j + 1]);
          }
          char* q = reinterpret_cast<char*>(reinterpret_cast<T*>(P.y) + ((size_t)n * P.hw + r0 + my_row) * P.c + ch);
          const size_t qstride = (size_t)rpb * P.c * sizeof(T);
#pragma unroll 4
          for (int r = my_row; r < nrows; r += rpb, q += qstride) {
            float2 v[E / 2];
            P16<T>::unpack(*reinterpret_cast<const uint4*>(my_slab + (size_t)r * ld), v);
#pragma unroll
            for (int j = 0; j < E / 2; ++j) {
              float2 t = __ffma2_rn(v[j], scale2[j], shift2[j]);
              if (P.silu) {
                float2 th;
                asm("tanh.approx.f32 %0, %1;" : "=f"(th.x) : "f"(t.x));
                asm("tanh.approx.f32 %0, %1;" : "=f"(th.y) : "f"(t.y));
                t = __ffma2_rn(t, th, t);
              }
              v[j] = t;
            }
            if (bulk_out) *reinterpret_cast<uint4*>(const_cast<T*>(my_slab) + (size_t)r * ld) = P16<T>::pack(v);     // in place
            else st_na_v4(q, P16<T>::pack(v));
          }
          if (bulk_out) sm100::fence_proxy_async();       // generic-proxy writes -> visible to the bulk store
        }
        gws_bar(2 + ag);                                  // every thread is done with the buffer
        if (gt == 0) {
          if (bulk_out) {
            // one source: the normalised slab is contiguous in y -- it leaves as bulk stores (the TMA engine's writes instead
            // of 16-byte stores from every thread); the buffer goes back once they have READ it
            const char* src = reinterpret_cast<const char*>(bufs + (size_t)b * kGwsBufBytes);
            char* dst = reinterpret_cast<char*>(reinterpret_cast<T*>(P.y) + ((size_t)n * P.hw + r0) * P.c);
            const uint32_t bytes = (uint32_t)((size_t)nrows * P.c * sizeof(T));
            for (uint32_t o = 0; o < bytes; o += 16384u)
              asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                           :: "l"(dst + o), "r"(sm100::smem_u32(src + o)), "r"(min(16384u, bytes - o)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
          }
          sm100::mbar_arrive(&H->empty[b]);
        }
      }
    }
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");       // bulk stores of this thread (if any) complete before exit
  // the last CTA out re-arms the set (ticket, exit count, arrival counters and ready flags of all samples)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    H->item[0] = (int)atomicAdd(&sync[1], 1u);
  }
  __syncthreads();
  if (H->item[0] == (int)gridDim.x - 1) {
    for (int i = threadIdx.x; i < P.n; i += blockDim.x) { sync[2 + 2 * i] = 0u; sync[3 + 2 * i] = 0u; }
    if (threadIdx.x == 0) { sync[0] = 0u; sync[1] = 0u; }
  }
}

static int gn_ws_launch(GnParams& P, int tpr, int rpb, cudaStream_t st) {
  static unsigned launch_no3 = 0;
  const int set = (int)(launch_no3++ % kGnSyncSets);
  const size_t smem = 1024 + kGwsGroup * sizeof(float4) + kGwsBufs * kGwsBufBytes;
  static bool attr = false;
  if (!attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(gn_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  long long grid = (long long)P.n * P.slabs;
  if (grid > num_sms()) grid = num_sms();
  static int nowait = -1;
  if (nowait < 0) { const char* e_ = getenv("VF_GN_DEBUG_NOWAIT"); nowait = e_ ? atoi(e_) : 0; }
  static int ws_mode = -1;           // VF_GN_WS=2: single-source slabs leave as bulk stores out of shared memory
  if (ws_mode < 0) { const char* e_ = getenv("VF_GN_WS"); ws_mode = e_ ? atoi(e_) : 0; }
  gn_ws_kernel<<<(int)grid, kGwsThreads, smem, st>>>(P, tpr, rpb, set | (nowait ? 0x100 : 0) | (ws_mode >= 2 ? 0x200 : 0));
  return check_cuda(cudaGetLastError(), "gn_ws_kernel launch");
}

// ================================================================================================
// residual add + bias + LayerNorm (one warp per row)
// ================================================================================================
constexpr int kLnWarps = 8;

struct LnParams {
  const void* x; const void* y; const void* bias; const void* gamma; const void* beta;
  void* res; void* out;
  long long rows, rows_per_bias;
  int c;
  float eps;
};

// R rows per warp, all of their loads issued before the first reduction: one row per warp (640 B at c = 320) left
// an SM with 41 KB in flight and the single-input form at 44 % of HBM.
template <typename T, int MAXC, int R>   // MAXC: 16-byte chunks per lane
__global__ void __launch_bounds__(kLnWarps * 32)
add_layer_norm_kernel(const LnParams P) {
  constexpr int E = V16<T>::E;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * R;
  if (row0 >= P.rows) return;
  const int chunks = P.c / E;
  float v[R][MAXC][E];
  float s[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r;
    s[r] = 0.f;
    if (row >= P.rows) continue;
    const T* x = reinterpret_cast<const T*>(P.x) + row * P.c;
    const T* y = P.y ? reinterpret_cast<const T*>(P.y) + row * P.c : nullptr;
    const T* bias = nullptr;
    if (P.bias) bias = reinterpret_cast<const T*>(P.bias) + (P.rows_per_bias > 0 ? (row / P.rows_per_bias) * P.c : 0);
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int ch = lane + 32 * k;
      if (ch < chunks) {
        V16<T>::ld(x + ch * E, v[r][k]);
        if (y) {
          float t[E];
          V16<T>::ld(y + ch * E, t);
#pragma unroll
          for (int j = 0; j < E; ++j) v[r][k][j] += t[j];
        }
        if (bias) {
          float t[E];
          V16<T>::ld(bias + ch * E, t);
#pragma unroll
          for (int j = 0; j < E; ++j) v[r][k][j] += t[j];
        }
        if (P.res) {
          // the residual stream is stored in T; normalise what was stored so both consumers agree
          V16<T>::st(reinterpret_cast<T*>(P.res) + row * P.c + ch * E, v[r][k]);
          if (sizeof(T) == 2) {
#pragma unroll
            for (int j = 0; j < E; ++j) v[r][k][j] = __bfloat162float(__float2bfloat16_rn(v[r][k][j]));
          }
        }
#pragma unroll
        for (int j = 0; j < E; ++j) s[r] += v[r][k][j];
      }
    }
  }
  const T* gamma = reinterpret_cast<const T*>(P.gamma);
  const T* beta = reinterpret_cast<const T*>(P.beta);
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r;
    if (row >= P.rows) continue;                       // warp-uniform
    float sum = s[r];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / (float)P.c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      if (lane + 32 * k < chunks) {
#pragma unroll
        for (int j = 0; j < E; ++j) { const float t = v[r][k][j] - mean; q = fmaf(t, t, q); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q / (float)P.c + P.eps);
    T* out = reinterpret_cast<T*>(P.out) + row * P.c;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int ch = lane + 32 * k;
      if (ch < chunks) {
        float g[E], b[E];
        V16<T>::ld(gamma + ch * E, g);
        V16<T>::ld(beta + ch * E, b);
#pragma unroll
        for (int j = 0; j < E; ++j) v[r][k][j] = fmaf((v[r][k][j] - mean) * rstd, g[j], b[j]);
        V16<T>::st(out + ch * E, v[r][k]);
      }
    }
  }
}

// Pure LayerNorm (no add, nothing to store back): the rows stay PACKED in registers (4 per 16-byte chunk instead of
// 8 floats), so R = 4 rows per warp cost ~60 registers and four CTAs stay resident: ~80 KB of loads in flight per SM.
template <typename T, int MAXC, int R>
__global__ void __launch_bounds__(kLnWarps * 32, MAXC <= 2 ? 4 : 3)
layer_norm_kernel(const LnParams P) {
  constexpr int E = V16<T>::E;
  const int lane = threadIdx.x & 31;
  const long long row0 = ((long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * R;
  if (row0 >= P.rows) return;
  const int chunks = P.c / E;
  uint4 raw[R][MAXC];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r < P.rows ? row0 + r : P.rows - 1;      // clamp: tail rows are recomputed, not stored
    const T* x = reinterpret_cast<const T*>(P.x) + row * P.c;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int ch = lane + 32 * k;
      raw[r][k] = ch < chunks ? ld_nc_v4(x + ch * E) : make_uint4(0u, 0u, 0u, 0u);
    }
  }
  const T* gamma = reinterpret_cast<const T*>(P.gamma);
  const T* beta = reinterpret_cast<const T*>(P.beta);
  const float inv_c = 1.0f / (float)P.c;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const long long row = row0 + r;
    if (row >= P.rows) break;                          // warp-uniform
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      float v[E];
      V16<T>::unpack(raw[r][k], v);                    // padding chunks are zero
#pragma unroll
      for (int j = 0; j < E; ++j) sum += v[j];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      if (lane + 32 * k < chunks) {
        float v[E];
        V16<T>::unpack(raw[r][k], v);
#pragma unroll
        for (int j = 0; j < E; ++j) { const float t = v[j] - mean; q = fmaf(t, t, q); }
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + P.eps);
    T* out = reinterpret_cast<T*>(P.out) + row * P.c;
#pragma unroll
    for (int k = 0; k < MAXC; ++k) {
      const int ch = lane + 32 * k;
      if (ch < chunks) {
        float v[E], g[E], b[E];
        V16<T>::unpack(raw[r][k], v);
        V16<T>::ld(gamma + ch * E, g);
        V16<T>::ld(beta + ch * E, b);
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = fmaf((v[j] - mean) * rstd, g[j], b[j]);
        st_na_v4(out + ch * E, V16<T>::pack(v));
      }
    }
  }
}

// Pure LayerNorm for c = 40 L chunks (L = 8, 16, 32: the UNet's 320 / 640 / 1280 channels in bf16): L lanes per
// row, five 16-byte chunks per lane, 32 / L rows side by side in a warp and kPasses such groups in flight.  The
// one-row-per-warp mapping above issues two chunk iterations per row at c = 320 with 24 of 32 lanes idle in the
// second and ten shuffles per row; ncu showed it ISSUE-bound (83 % issue-active at 47 % of DRAM).  Here every lane
// works in every iteration and a 3-step (L = 8) shuffle serves four rows at once.
template <typename T, int L, int kPasses>
__global__ void __launch_bounds__(kLnWarps * 32, 3)
layer_norm_rows_kernel(const LnParams P) {
  constexpr int E = V16<T>::E;
  constexpr int CH = 5;
  constexpr int kRowsPerPass = 32 / L;
  const int lane = threadIdx.x & 31;
  const int sub = lane / L, l = lane % L;                     // row slot within a pass, lane within the row
  const long long row0 = ((long long)blockIdx.x * kLnWarps + (threadIdx.x >> 5)) * (kRowsPerPass * kPasses);
  if (row0 >= P.rows) return;
  uint4 raw[kPasses][CH];
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    long long row = row0 + p * kRowsPerPass + sub;
    if (row >= P.rows) row = P.rows - 1;                       // clamp: recomputed, not stored
    const T* x = reinterpret_cast<const T*>(P.x) + row * P.c;
#pragma unroll
    for (int k = 0; k < CH; ++k) raw[p][k] = ld_nc_v4(x + (l + L * k) * E);
  }
  const T* gamma = reinterpret_cast<const T*>(P.gamma);
  const T* beta = reinterpret_cast<const T*>(P.beta);
  const float inv_c = 1.0f / (float)P.c;
#pragma unroll
  for (int p = 0; p < kPasses; ++p) {
    const long long row = row0 + p * kRowsPerPass + sub;
    float sum = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      float v[E];
      V16<T>::unpack(raw[p][k], v);
#pragma unroll
      for (int j = 0; j < E; ++j) sum += v[j];
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float q = 0.f;
#pragma unroll
    for (int k = 0; k < CH; ++k) {
      float v[E];
      V16<T>::unpack(raw[p][k], v);
#pragma unroll
      for (int j = 0; j < E; ++j) { const float t = v[j] - mean; q = fmaf(t, t, q); }
    }
#pragma unroll
    for (int o = L / 2; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_c + P.eps);
    const float nmr = -mean * rstd;
    if (row < P.rows) {
      T* out = reinterpret_cast<T*>(P.out) + row * P.c;
#pragma unroll
      for (int k = 0; k < CH; ++k) {
        float v[E], g[E], b[E];
        V16<T>::unpack(raw[p][k], v);
        V16<T>::ld(gamma + (l + L * k) * E, g);              // 640 B .. 2.5 KB per CTA: L1 hits after the first pass
        V16<T>::ld(beta + (l + L * k) * E, b);
#pragma unroll
        for (int j = 0; j < E; ++j) v[j] = fmaf(fmaf(v[j], rstd, nmr), g[j], b[j]);     // ((v - mean) rstd) g + b
        st_na_v4(out + (l + L * k) * E, V16<T>::pack(v));
      }
    }
  }
}

// ================================================================================================
// GEGLU and residual add
// ================================================================================================
// exact (erf) GELU.  fp32 path: libdevice erff.  bf16 path: Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far
// below the 4e-3 spacing of the bf16 result) -- one MUFU.RCP, one MUFU.EX2 and a degree-5 Horner chain instead
// of erff's ~30 instructions, which made the kernel ALU-bound (48 bytes per 8 erff calls).
template <typename T> __device__ __forceinline__ float gelu_f(float g);
template <> __device__ __forceinline__ float gelu_f<float>(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752f)); }
template <> __device__ __forceinline__ float gelu_f<__nv_bfloat16>(float g) {
  const float x = fabsf(g) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, x, 1.0f));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float e = p * t * exp2f(-1.4426950408889634f * x * x);      // 1 - erf(|x|)
  const float erf_abs = 1.0f - e;
  return 0.5f * g * (1.0f + copysignf(erf_abs, g));
}

template <typename T>
__global__ void __launch_bounds__(256)
geglu_kernel(const T* __restrict__ h, T* __restrict__ out, long long rows, int k, long long ld_h) {
  constexpr int E = V16<T>::E;
  const int chunks = k / E;
  const long long total = rows * chunks;
  const long long stride = (long long)gridDim.x * blockDim.x;
  // two independent chunks per trip (four 16-byte loads in flight per thread)
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += 2 * stride) {
    const long long i1 = i + stride;
    const bool two = i1 < total;
    long long r0, r1;
    if (total < 0x7fffffffLL) {        // 32-bit index arithmetic whenever it fits (64-bit division is ~40 instructions)
      r0 = (unsigned)i / (unsigned)chunks;
      r1 = two ? (unsigned)i1 / (unsigned)chunks : r0;
    } else {
      r0 = i / chunks;
      r1 = two ? i1 / chunks : r0;
    }
    const int c0 = (int)(i - r0 * chunks) * E;
    const int c1 = two ? (int)(i1 - r1 * chunks) * E : c0;
    const uint4 ua0 = ld_nc_v4(h + r0 * ld_h + c0), ug0 = ld_nc_v4(h + r0 * ld_h + k + c0);
    const uint4 ua1 = ld_nc_v4(h + r1 * ld_h + c1), ug1 = ld_nc_v4(h + r1 * ld_h + k + c1);
    float a[E], g[E];
    V16<T>::unpack(ua0, a);
    V16<T>::unpack(ug0, g);
#pragma unroll
    for (int j = 0; j < E; ++j) a[j] *= gelu_f<T>(g[j]);
    st_na_v4(out + r0 * (long long)k + c0, V16<T>::pack(a));
    if (two) {
      V16<T>::unpack(ua1, a);
      V16<T>::unpack(ug1, g);
#pragma unroll
      for (int j = 0; j < E; ++j) a[j] *= gelu_f<T>(g[j]);
      st_na_v4(out + r1 * (long long)k + c1, V16<T>::pack(a));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
add_bias_kernel(const T* a, const T* __restrict__ b, const T* __restrict__ bias, T* out,     // out may alias a (in place)
                long long rows, int c, long long rows_per_bias) {
  constexpr int E = V16<T>::E;
  const int chunks = c / E;
  const long long total = rows * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / chunks;
    const int ch = (int)(i - r * chunks) * E;
    float va[E];
    V16<T>::ld(a + r * c + ch, va);
    if (b) {
      float vb[E];
      V16<T>::ld(b + r * c + ch, vb);
#pragma unroll
      for (int j = 0; j < E; ++j) va[j] += vb[j];
    }
    if (bias) {
      float vc[E];
      V16<T>::ld(bias + (rows_per_bias > 0 ? (r / rows_per_bias) * c : 0) + ch, vc);
#pragma unroll
      for (int j = 0; j < E; ++j) va[j] += vc[j];
    }
    V16<T>::st(out + r * c + ch, va);
  }
}

// nearest-neighbour 2x upsampling of a channels-last map: out[n, 2y+a, 2x+b, :] = in[n, y, x, :].
// (ATen's upsample_nearest2d NHWC kernel moves one element per thread: 0.8 ms for the 32x32x640 -> 64x64 level
// of a 96-sample batch, 8x the time the 0.63 GB of traffic needs.)  One thread = one 16-byte chunk of an INPUT
// pixel, stored four times.
template <typename T>
__global__ void __launch_bounds__(256)
upsample2x_kernel(const T* __restrict__ in, T* __restrict__ out, long long n_pix, int h, int w, int chunks) {
  constexpr int E = V16<T>::E;
  const long long total = n_pix * chunks;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long pix = i / chunks;
    const int ch = (int)(i - pix * chunks) * E;
    const int x = (int)(pix % w);
    const long long t = pix / w;
    const int y = (int)(t % h);
    const long long n = t / h;
    const uint4 v = ld_nc_v4(in + pix * (long long)chunks * E + ch);
    T* o = out + (((n * 2 * h + 2 * y) * 2 * w) + 2 * x) * (long long)chunks * E + ch;
    const long long row = 2LL * w * chunks * E;
    st_na_v4(o, v);
    st_na_v4(o + (long long)chunks * E, v);
    st_na_v4(o + row, v);
    st_na_v4(o + row + (long long)chunks * E, v);
  }
}

static int ew_grid(long long total) {
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vf

extern "C" long long vf_group_norm_workspace_floats(int n, int hw, int groups) {
  int slabs, rps;
  if (n <= 0 || hw <= 0 || groups <= 0) return 0;
  vf::gn_plan(n, hw, &slabs, &rps);
  (void)slabs;
  return (long long)n * 129 * groups * 2;     // 128 = the most slabs any of the slab plans hands out; + per-sample totals (gn_ws_kernel)
}

extern "C" int vf_group_norm_nhwc(const void* x, const void* add_nc, const void* gamma, const void* beta, void* y,
                                  float* workspace, int n, int hw, int c, int groups, float eps, int silu,
                                  int dtype, void* stream) {
  return vf_group_norm_nhwc_cat(x, c, nullptr, 0, add_nc, gamma, beta, y, workspace, n, hw, groups, eps, silu, dtype, stream);
}

extern "C" int vf_group_norm_nhwc_cat(const void* x, int c1, const void* x2, int c2, const void* add_nc,
                                      const void* gamma, const void* beta, void* y, float* workspace,
                                      int n, int hw, int groups, float eps, int silu, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  const int c = c1 + (x2 ? c2 : 0);
  if (!x || !gamma || !beta || !y || !workspace) return fail("vf_group_norm_nhwc: null pointer");
  if (x2 && (c2 <= 0 || c1 % (dtype == VF_F32 ? 4 : 8) || c2 % (dtype == VF_F32 ? 4 : 8) || !al16(x2)))
    return fail("vf_group_norm_nhwc_cat: both channel counts must be multiples of a 16-byte vector (c1=%d c2=%d)", c1, c2);
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("vf_group_norm_nhwc: bad dtype %d", dtype);
  const int e = dtype == VF_F32 ? 4 : 8;
  if (n <= 0 || hw <= 0 || c <= 0 || groups <= 0 || groups > kGnMaxGroups || c % groups || c % e)
    return fail("vf_group_norm_nhwc: bad shape n=%d hw=%d c=%d groups=%d", n, hw, c, groups);
  if (c / e > 3 * kGnThreads) return fail("vf_group_norm_nhwc: c=%d too wide", c);
  if (n > 65535) return fail("vf_group_norm_nhwc: n=%d exceeds 65535", n);
  if (!al16(x) || !al16(y) || !al16(gamma) || !al16(beta) || (add_nc && !al16(add_nc)))
    return fail("vf_group_norm_nhwc: pointers must be 16-byte aligned");
  GnParams P;
  P.x = x; P.add_nc = add_nc; P.gamma = gamma; P.beta = beta; P.y = y; P.ws = workspace;
  P.x2 = x2; P.c1 = x2 ? c1 : c;
  P.n = n; P.hw = hw; P.c = c; P.groups = groups; P.eps = eps; P.silu = silu;
  gn_plan(n, hw, &P.slabs, &P.rows_per_slab);
  cudaStream_t st = (cudaStream_t)stream;
  const int cpt = (c / e + kGnThreads - 1) / kGnThreads;
  const int chunks_ = c / e;
  const int rpb_ = kGnThreads / (chunks_ < kGnThreads ? chunks_ : kGnThreads);
  const size_t stats_smem = (size_t)2 * rpb_ * c * sizeof(float);
  if (stats_smem > 48 * 1024) return fail("vf_group_norm_nhwc: c=%d too wide", c);
  {
    // Default (VF_GN_FUSED=2): one persistent launch with the slab resident in shared memory; 1 = one launch, second
    // read out of L2; 0 = the statistics / apply pair.
    static int fused = -1, slab_kb = -1;
    if (fused < 0) { const char* e_ = getenv("VF_GN_FUSED"); fused = e_ ? atoi(e_) : 2; }
    if (slab_kb < 0) { const char* e_ = getenv("VF_GN_SLAB_KB"); slab_kb = e_ ? atoi(e_) : 160; if (slab_kb < 1) slab_kb = 160; }
    if (fused && chunks_ <= kGnfMaxThreads && n <= kGnfMaxN) {
      const int tpr = chunks_, rpb = kGnfMaxThreads / tpr;
      const int threads = (tpr * rpb + 31) / 32 * 32;
      // fused >= 2 (default): slab resident in shared memory (x crosses HBM once); 1: second read out of L2
      if (fused >= 2 && c / groups >= e) {
        static int res_threads = -1;                    // VF_GN_RES_THREADS: most threads per CTA (multiple of 32, <= 512)
        if (res_threads < 0) { const char* e_ = getenv("VF_GN_RES_THREADS"); res_threads = e_ ? atoi(e_) : kGnrMaxThreads; if (res_threads < 32 || res_threads > kGnrMaxThreads) res_threads = kGnrMaxThreads; }
        const int mt = res_threads < tpr ? tpr : res_threads;
        const int rpb_r = mt / tpr;
        const int threads_r = (tpr * rpb_r + 31) / 32 * 32;
        const size_t budget = gnr_smem_per_cta() - kGnrHeaderBytes - kGnrMaxThreads * sizeof(float4);
        // VF_GN_WS=1: the warp-specialised form (gn_ws_kernel; bf16, rows of at most 256 16-byte chunks, enough items to give
        // every SM two)
        static int ws_knob = -1;
        if (ws_knob < 0) { const char* e_ = getenv("VF_GN_WS"); ws_knob = e_ ? atoi(e_) : 0; }
        if (ws_knob && dtype == VF_BF16 && tpr <= kGwsGroup) {
          const int rpb_w = kGwsGroup / tpr;
          gnr_plan(n, hw, c, 2, rpb_w, kGwsBufBytes, &P.slabs, &P.rows_per_slab);
          if (P.slabs > 0 && P.slabs <= kGwsSlabStride && (long long)n * P.slabs >= 2LL * num_sms())
            return gn_ws_launch(P, tpr, rpb_w, st);
        }
        // VF_GN_PIPE=1 (opt-in): two half-size slabs per CTA, the front half of the next item ahead of the wait of the
        // current one (gn_resident2_kernel).  Correct, and SLOWER (0.180 vs 0.133 ms at 96 x 4096 x 320; the same with the
        // wait skipped): what an item costs is its fixed chain of block barriers, reductions and L2 round trips, and half-size
        // slabs pay it twice as often (profiles/r2_gn_pipe_ab.txt).  Default 0: one ~100 KB slab per CTA.
        static int pipe = -1;
        if (pipe < 0) { const char* e_ = getenv("VF_GN_PIPE"); pipe = e_ ? atoi(e_) : 0; }
        if (pipe) {
          gnr_plan(n, hw, c, dtype == VF_F32 ? 4 : 2, rpb_r, budget / 2, &P.slabs, &P.rows_per_slab);
          if (P.slabs > 0 && (long long)n * P.slabs >= 4LL * num_sms())
            return dtype == VF_F32 ? gn_resident2_launch<float>(P, tpr, rpb_r, threads_r, st)
                                   : gn_resident2_launch<__nv_bfloat16>(P, tpr, rpb_r, threads_r, st);
        }
        gnr_plan(n, hw, c, dtype == VF_F32 ? 4 : 2, rpb_r, budget, &P.slabs, &P.rows_per_slab);
        if (P.slabs > 0)
          return dtype == VF_F32 ? gn_resident_launch<float>(P, tpr, rpb_r, threads_r, st)
                                 : gn_resident_launch<__nv_bfloat16>(P, tpr, rpb_r, threads_r, st);
      }
      gnf_plan(hw, c, dtype == VF_F32 ? 4 : 2, rpb, slab_kb, &P.slabs, &P.rows_per_slab);
      return dtype == VF_F32 ? gn_fused_launch<float>(P, tpr, rpb, threads, st)
                             : gn_fused_launch<__nv_bfloat16>(P, tpr, rpb, threads, st);
    }
  }
  {
    // Experimental (VF_GN_ONEPASS=1): measured SLOWER than the L2-blocked two-pass below on B200 (0.49 vs
    // 0.24 ms at n=96, 64x64x320 bf16): one 256-thread CTA per SM cannot hide its own shared-memory and MUFU
    // latencies, and load / statistics / apply phases of a CTA do not overlap.
    static int one_pass = -1;
    if (one_pass < 0) { const char* e_ = getenv("VF_GN_ONEPASS"); one_pass = e_ ? atoi(e_) : 0; }
    size_t cl_smem = 0;
    const int cs = one_pass ? gn_cluster_plan(n, hw, c, dtype == VF_F32 ? 4 : 2, rpb_, &cl_smem) : 0;
    if (cs > 0) {
      P.rows_per_slab = (hw + cs - 1) / cs;
      P.slabs = cs;
#define VF_GNC(T)                                                                                 \
      (cpt == 1 ? gn_cluster_launch<T, 1>(P, cs, cl_smem, st)                                      \
                : cpt == 2 ? gn_cluster_launch<T, 2>(P, cs, cl_smem, st) : gn_cluster_launch<T, 3>(P, cs, cl_smem, st))
      const int rc_cl = dtype == VF_F32 ? VF_GNC(float) : VF_GNC(__nv_bfloat16);
#undef VF_GNC
      if (rc_cl >= 0) return rc_cl;
      gn_plan(n, hw, &P.slabs, &P.rows_per_slab);      // cluster shape not schedulable here: two-pass
    }
  }
  // Experimental L2 blocking (VF_GN_L2_MB=<MiB>, default off): statistics and apply back to back over groups
  // of samples small enough to stay in the 126 MB L2 between the passes.  Measured SLOWER on B200 (0.32-0.39 vs
  // 0.24 ms at n=96, 64x64x320 bf16): six small launch pairs lose more to tails than the L2 hits return.
  static long long l2_bytes = -1;
  if (l2_bytes < 0) { const char* e_ = getenv("VF_GN_L2_MB"); l2_bytes = (e_ ? atoll(e_) : 0) << 20; }
  const size_t per_sample = (size_t)hw * c * (dtype == VF_F32 ? 4 : 2);
  int group = n;
  if (l2_bytes > 0 && per_sample * n > (size_t)l2_bytes) {
    group = (int)((size_t)l2_bytes / per_sample);
    if (group < 1) group = 1;
  }
  const int c2_ = x2 ? c - P.c1 : 0;
  const size_t es = dtype == VF_F32 ? 4 : 2;
#define VF_GN_LAUNCH(T, K)                                        \
  do {                                                            \
    gn_stats_kernel<T, K><<<grid, kGnThreads, stats_smem, st>>>(P);        \
    gn_apply_kernel<T, K><<<grid, kGnThreads, 0, st>>>(P);        \
  } while (0)
  for (int n0 = 0; n0 < n; n0 += group) {
    const int gn = n - n0 < group ? n - n0 : group;
    P.n = gn;
    P.x = static_cast<const char*>(x) + (size_t)n0 * hw * P.c1 * es;
    P.x2 = x2 ? static_cast<const char*>(x2) + (size_t)n0 * hw * c2_ * es : nullptr;
    P.y = static_cast<char*>(y) + (size_t)n0 * hw * c * es;
    P.add_nc = add_nc ? static_cast<const char*>(add_nc) + (size_t)n0 * c * es : nullptr;
    gn_plan(gn, hw, &P.slabs, &P.rows_per_slab);
    P.ws = workspace;       // reused: groups run in stream order
    const dim3 grid(P.slabs, gn);
    if (dtype == VF_F32) {
      if (cpt == 1) VF_GN_LAUNCH(float, 1); else if (cpt == 2) VF_GN_LAUNCH(float, 2); else VF_GN_LAUNCH(float, 3);
    } else {
      if (cpt == 1) VF_GN_LAUNCH(__nv_bfloat16, 1); else if (cpt == 2) VF_GN_LAUNCH(__nv_bfloat16, 2); else VF_GN_LAUNCH(__nv_bfloat16, 3);
    }
  }
#undef VF_GN_LAUNCH
  return check_cuda(cudaGetLastError(), "group_norm kernels launch");
}

extern "C" int vf_add_layer_norm(const void* x, const void* y, const void* bias, long long rows_per_bias,
                                 const void* gamma, const void* beta, void* res, void* out,
                                 long long rows, int c, float eps, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !gamma || !beta || !out) return fail("vf_add_layer_norm: null pointer");
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("vf_add_layer_norm: bad dtype %d", dtype);
  const int e = dtype == VF_F32 ? 4 : 8;
  if (rows <= 0 || c <= 0 || c % e) return fail("vf_add_layer_norm: bad shape rows=%lld c=%d", rows, c);
  const int per_lane = (c / e + 31) / 32;
  if (per_lane > 10) return fail("vf_add_layer_norm: c=%d too wide", c);
  if ((y || bias) && !res) return fail("vf_add_layer_norm: res is required when y or bias is given");
  if (!al16(x) || !al16(out) || !al16(gamma) || !al16(beta) || (y && !al16(y)) || (bias && !al16(bias)) || (res && !al16(res)))
    return fail("vf_add_layer_norm: pointers must be 16-byte aligned");
  LnParams P;
  P.x = x; P.y = y; P.bias = bias; P.gamma = gamma; P.beta = beta; P.res = res; P.out = out;
  P.rows = rows; P.rows_per_bias = rows_per_bias; P.c = c; P.eps = eps;
  const bool pure = !y && !bias && !res;
  // rows per warp: only the pure form (packed registers) gains from several rows in flight; with the add the extra
  // registers cost more occupancy than the rows bring (measured 0.24 -> 0.43 ms at c = 320), so it stays at one
  const int rpw = !pure ? 1 : per_lane <= 2 ? 4 : per_lane <= 5 ? 2 : 1;
  const long long blocks = (rows + (long long)kLnWarps * rpw - 1) / ((long long)kLnWarps * rpw);
  if (blocks > 2147483647LL) return fail("vf_add_layer_norm: too many rows");
  cudaStream_t st = (cudaStream_t)stream;
  const int thr = kLnWarps * 32;
  // the UNet's widths (40 L chunks of 16 bytes, L = 8 / 16 / 32): L lanes per row, several rows per warp
  const int chunks_total = c / e;
  const int lanes_per_row = (pure && chunks_total % 5 == 0) ? chunks_total / 5 : 0;
  if (lanes_per_row == 8 || lanes_per_row == 16 || lanes_per_row == 32) {
    const int passes = lanes_per_row == 32 ? 2 : lanes_per_row == 16 ? 2 : 2;
    const long long rows_per_warp = (32 / lanes_per_row) * passes;
    const long long nb = (rows + kLnWarps * rows_per_warp - 1) / (kLnWarps * rows_per_warp);
    if (nb > 2147483647LL) return fail("vf_add_layer_norm: too many rows");
#define VF_LNR_LAUNCH(T)                                                                              \
    do {                                                                                               \
      if (lanes_per_row == 8) layer_norm_rows_kernel<T, 8, 2><<<(int)nb, thr, 0, st>>>(P);              \
      else if (lanes_per_row == 16) layer_norm_rows_kernel<T, 16, 2><<<(int)nb, thr, 0, st>>>(P);       \
      else layer_norm_rows_kernel<T, 32, 2><<<(int)nb, thr, 0, st>>>(P);                                \
    } while (0)
    if (dtype == VF_F32) VF_LNR_LAUNCH(float); else VF_LNR_LAUNCH(__nv_bfloat16);
#undef VF_LNR_LAUNCH
    return check_cuda(cudaGetLastError(), "layer_norm_rows_kernel launch");
  }
#define VF_LN_LAUNCH(T)                                                                      \
  do {                                                                                       \
    if (pure && per_lane <= 2) layer_norm_kernel<T, 2, 4><<<(int)blocks, thr, 0, st>>>(P);    \
    else if (pure && per_lane <= 5) layer_norm_kernel<T, 5, 2><<<(int)blocks, thr, 0, st>>>(P); \
    else if (per_lane <= 2) add_layer_norm_kernel<T, 2, 1><<<(int)blocks, thr, 0, st>>>(P);   \
    else if (per_lane <= 5) add_layer_norm_kernel<T, 5, 1><<<(int)blocks, thr, 0, st>>>(P);   \
    else add_layer_norm_kernel<T, 10, 1><<<(int)blocks, thr, 0, st>>>(P);                     \
  } while (0)
  if (dtype == VF_F32) VF_LN_LAUNCH(float); else VF_LN_LAUNCH(__nv_bfloat16);
#undef VF_LN_LAUNCH
  return check_cuda(cudaGetLastError(), "add_layer_norm_kernel launch");
}

extern "C" int vf_geglu(const void* h, void* out, long long rows, int k, long long ld_h, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!h || !out) return fail("vf_geglu: null pointer");
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("vf_geglu: bad dtype %d", dtype);
  const int e = dtype == VF_F32 ? 4 : 8;
  if (rows <= 0 || k <= 0 || k % e || ld_h < 2LL * k || ld_h % e) return fail("vf_geglu: bad shape rows=%lld k=%d ld=%lld", rows, k, ld_h);
  if (!al16(h) || !al16(out)) return fail("vf_geglu: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(rows * (k / e));
  if (dtype == VF_F32) geglu_kernel<float><<<grid, 256, 0, st>>>((const float*)h, (float*)out, rows, k, ld_h);
  else geglu_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)h, (__nv_bfloat16*)out, rows, k, ld_h);
  return check_cuda(cudaGetLastError(), "geglu_kernel launch");
}

extern "C" int vf_add_bias(const void* a, const void* b, const void* bias, long long rows_per_bias, void* out,
                           long long rows, int c, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!a || !out) return fail("vf_add_bias: null pointer");
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("vf_add_bias: bad dtype %d", dtype);
  const int e = dtype == VF_F32 ? 4 : 8;
  if (rows <= 0 || c <= 0 || c % e) return fail("vf_add_bias: bad shape rows=%lld c=%d", rows, c);
  if (!al16(a) || !al16(out) || (b && !al16(b)) || (bias && !al16(bias))) return fail("vf_add_bias: pointers must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = ew_grid(rows * (c / e));
  if (dtype == VF_F32)
    add_bias_kernel<float><<<grid, 256, 0, st>>>((const float*)a, (const float*)b, (const float*)bias, (float*)out, rows, c, rows_per_bias);
  else
    add_bias_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b, (const __nv_bfloat16*)bias,
                                                          (__nv_bfloat16*)out, rows, c, rows_per_bias);
  return check_cuda(cudaGetLastError(), "add_bias_kernel launch");
}

extern "C" int vf_upsample_nearest2x_nhwc(const void* x, void* out, int n, int h, int w, int c, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !out) return fail("vf_upsample_nearest2x_nhwc: null pointer");
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("vf_upsample_nearest2x_nhwc: bad dtype %d", dtype);
  const int e = dtype == VF_F32 ? 4 : 8;
  if (n <= 0 || h <= 0 || w <= 0 || c <= 0 || c % e) return fail("vf_upsample_nearest2x_nhwc: bad shape n=%d h=%d w=%d c=%d", n, h, w, c);
  if (!al16(x) || !al16(out)) return fail("vf_upsample_nearest2x_nhwc: pointers must be 16-byte aligned");
  const long long n_pix = (long long)n * h * w;
  const int grid = ew_grid(n_pix * (c / e));
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VF_F32) upsample2x_kernel<float><<<grid, 256, 0, st>>>((const float*)x, (float*)out, n_pix, h, w, c / e);
  else upsample2x_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)out, n_pix, h, w, c / e);
  return check_cuda(cudaGetLastError(), "upsample2x_kernel launch");
}
