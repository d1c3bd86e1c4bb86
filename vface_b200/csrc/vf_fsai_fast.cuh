// FSAI fast path: register-resident FFT with ONE shared-memory exchange per direction.
//
// (scripts/face_swap_utils.py:425-464; maths in vf_fsai.cu's header: out = dst + Re/Im ifft(h * fft(z)),
//  z = (donor - dst_a) + i (donor - dst_b).)
//
// A row of D = V*L*M channels is read as D/V vectors of V consecutive channels (one 8/16-byte access), so
// a thread holds V interleaved sub-sequences x[V*m + e], e < V, as the lanes of a CVec<V>.  The length-D
// transform is   sub-DFT of length L*M over m (same butterflies for all lanes)  +  radix-V combine over e:
//
//   phase A  (M threads per row, thread t):  m = t + M*r, r < L.  L-point DFT over r in registers,
//            twiddle W_{LM}^{t*k1}, store (k1, t) to shared memory.
//   phase B  (L threads per row, thread k1): M-point DFT over t in registers -> Y_e[k1 + L*k2];
//            twiddle W_D^{e*k'} and V-point DFT over the lanes -> bins k' + (D/V)*j; multiply by the
//            filter response; the same steps backwards (inverse V-point, conj twiddle, inverse M-point);
//            store back to the SAME shared-memory unit (private to the thread: no barrier in between).
//   phase C  (thread t again): conj twiddle, inverse L-point DFT, out = dst + y, vectorised store.
//
// Shared memory: one "unit" per (row, k1) = M*V complex + 16 B pad (stride 4*odd words), laid out
// [lane pair][t][re pair, im pair], so phase A/C move 16 B per (k1, lane pair) with the quarter-warp
// contiguous, and phase B's 16-B accesses of consecutive threads fall in distinct bank groups.
// Per row and direction: D*8 B written + D*8 B read -- against 8 stages of that in the Stockham kernel.
//
// The phase bodies are __host__ __device__ and take the thread index as an argument: the CPU check
// (tests/csrc/fsai_host_check.cu) runs them thread by thread against a naive DFT.
#pragma once

#include <cuda_bf16.h>
#include <cstdint>

#include "vf_fft_reg.cuh"

namespace vf {
namespace fsaifast {

using fftreg::CVec;
using fftreg::RVec;

struct Args {
  const void* donor;
  const void* dst_a;
  void* out_a;
  const void* dst_b;   // fused only
  void* out_b;
  long long rows;      // token rows
  long long n_pairs;   // complex rows: rows (fused) or ceil(rows/2)
  long long ld_donor, ld_a, ld_out_a, ld_b, ld_out_b;
  int split;
  int fused;
  int prefetch;        // L2-prefetch the next iteration's rows
};

template <typename T_, int D_, int V_, int L_, int M_, int RP_>
struct Cfg {
  using T = T_;
  static constexpr int D = D_, V = V_, L = L_, M = M_, RP = RP_;
  static_assert(D == V * L * M, "D = V*L*M");
  static constexpr int kUnitFloats = V * M * 2 + 4;            // + 16 B pad: stride/4 words is odd
  static_assert(((kUnitFloats / 4) & 1) == 1, "unit stride must be an odd number of 16-byte groups");
  static constexpr int kThreadsA = RP * M;
  static constexpr int kThreadsB = RP * L;
  static constexpr int kThreads = ((kThreadsA > kThreadsB ? kThreadsA : kThreadsB) + 31) / 32 * 32;
  static constexpr int kRecFloats = V == 4 ? 12 : 4;           // phase-B table record per sub-bin k'
  static constexpr int kSub = D / V;                           // sub-DFT length
  // shared memory (floats): exchange units | tw1[L][M] complex | rec[kSub][kRecFloats]
  static constexpr int kExchFloats = RP * L * kUnitFloats;
  static constexpr int kTw1Floats = L * M * 2;
  static constexpr int kRecTableFloats = kSub * kRecFloats;
  static constexpr int kSmemFloats = kExchFloats + kTw1Floats + kRecTableFloats;
};

// ---- V consecutive elements <-> floats ----------------------------------------------------------------
template <typename T, int V> struct Vld;
template <> struct Vld<float, 4> {
  typedef float4 Raw;
  static VF_HD Raw ld_raw(const float* p) { return *reinterpret_cast<const float4*>(p); }
  static VF_HD void unpack(const Raw& u, float (&v)[4]) { v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w; }
  static VF_HD void ld(const float* p, float (&v)[4]) {
    const float4 u = *reinterpret_cast<const float4*>(p);
    v[0] = u.x; v[1] = u.y; v[2] = u.z; v[3] = u.w;
  }
  static VF_HD void st(float* p, const float (&v)[4]) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
};
template <> struct Vld<float, 2> {
  typedef float2 Raw;
  static VF_HD Raw ld_raw(const float* p) { return *reinterpret_cast<const float2*>(p); }
  static VF_HD void unpack(const Raw& u, float (&v)[2]) { v[0] = u.x; v[1] = u.y; }
  static VF_HD void ld(const float* p, float (&v)[2]) {
    const float2 u = *reinterpret_cast<const float2*>(p);
    v[0] = u.x; v[1] = u.y;
  }
  static VF_HD void st(float* p, const float (&v)[2]) { *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]); }
};
VF_HD float bf_lo(uint32_t u) {
  union { uint32_t i; float f; } c; c.i = u << 16; return c.f;
}
VF_HD float bf_hi(uint32_t u) {
  union { uint32_t i; float f; } c; c.i = u & 0xffff0000u; return c.f;
}
VF_HD uint32_t bf_pack(float lo, float hi) {
  const __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  union { __nv_bfloat162 b; uint32_t i; } c; c.b = t; return c.i;
}
template <> struct Vld<__nv_bfloat16, 4> {
  typedef uint2 Raw;
  static VF_HD Raw ld_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint2*>(p); }
  static VF_HD void unpack(const Raw& u, float (&v)[4]) { v[0] = bf_lo(u.x); v[1] = bf_hi(u.x); v[2] = bf_lo(u.y); v[3] = bf_hi(u.y); }
  static VF_HD void ld(const __nv_bfloat16* p, float (&v)[4]) {
    const uint2 u = *reinterpret_cast<const uint2*>(p);
    v[0] = bf_lo(u.x); v[1] = bf_hi(u.x); v[2] = bf_lo(u.y); v[3] = bf_hi(u.y);
  }
  static VF_HD void st(__nv_bfloat16* p, const float (&v)[4]) {
    *reinterpret_cast<uint2*>(p) = make_uint2(bf_pack(v[0], v[1]), bf_pack(v[2], v[3]));
  }
};
template <> struct Vld<__nv_bfloat16, 2> {
  typedef uint32_t Raw;
  static VF_HD Raw ld_raw(const __nv_bfloat16* p) { return *reinterpret_cast<const uint32_t*>(p); }
  static VF_HD void unpack(const Raw& u, float (&v)[2]) { v[0] = bf_lo(u); v[1] = bf_hi(u); }
  static VF_HD void ld(const __nv_bfloat16* p, float (&v)[2]) {
    const uint32_t u = *reinterpret_cast<const uint32_t*>(p);
    v[0] = bf_lo(u); v[1] = bf_hi(u);
  }
  static VF_HD void st(__nv_bfloat16* p, const float (&v)[2]) { *reinterpret_cast<uint32_t*>(p) = bf_pack(v[0], v[1]); }
};

// ---- tables (built once per CTA) ------------------------------------------------------------------------
// tw1[k1][t] = W_{LM}^{t*k1};  rec[k'] = { h[k' + (D/V) j]/D, j < V ; W_D^{e k'}, 1 <= e < V }.
// `pre` is W_D^j for j < D as (cos, -sin), computed by the caller (sincospif on the device).
template <typename C>
VF_HD void build_tables(int idx, float* sm_tw1, float* sm_rec, const float2* pre, int split) {
  constexpr int D = C::D, V = C::V, L = C::L, M = C::M;
  if (idx < L * M) {
    const int k1 = idx / M, t = idx - k1 * M;
    const float2 w = pre[(V * t * k1) % D];
    sm_tw1[2 * idx] = w.x;
    sm_tw1[2 * idx + 1] = w.y;
  }
  if (idx < C::kSub) {
    float* r = sm_rec + idx * C::kRecFloats;
    for (int j = 0; j < V; ++j) {
      const int k = idx + C::kSub * j;
      const float hv = 0.5f * ((k >= split ? 1.0f : 0.0f) + ((((D - k) % D) >= split) ? 1.0f : 0.0f));
      r[j] = hv / (float)D;
    }
    for (int e = 1; e < V; ++e) {
      const float2 w = pre[(e * idx) % D];
      r[V + 2 * (e - 1)] = w.x;
      r[V + 2 * (e - 1) + 1] = w.y;
    }
  }
}

template <int V> VF_HD void load4(const float* p, float (&q)[4]) {
  const float4 u = *reinterpret_cast<const float4*>(p);
  q[0] = u.x; q[1] = u.y; q[2] = u.z; q[3] = u.w;
}
VF_HD void store4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }

// unit layout: [lane pair][t][re of the pair, im of the pair]: a 16-byte access is exactly the two
// 64-bit register pairs the packed arithmetic works on (no register shuffling on either side)
template <typename C>
VF_HD void unit_store(float* unit, int t, const CVec<C::V>& x) {
#pragma unroll
  for (int ep = 0; ep < C::V / 2; ++ep)
    store4(unit + ep * (C::M * 4) + t * 4, x.re.p[ep].x, x.re.p[ep].y, x.im.p[ep].x, x.im.p[ep].y);
}
template <typename C>
VF_HD void unit_load(const float* unit, int t, CVec<C::V>& x) {
#pragma unroll
  for (int ep = 0; ep < C::V / 2; ++ep) {
    float q[4];
    load4<C::V>(unit + ep * (C::M * 4) + t * 4, q);
    x.re.p[ep] = make_float2(q[0], q[1]);
    x.im.p[ep] = make_float2(q[2], q[3]);
  }
}

// ---- phase A: load, L-point DFT, twiddle, store -----------------------------------------------------------
template <typename C>
VF_HD void phase_a(const Args& A, long long pair_base, int tid, float* sm_exch, const float* sm_tw1) {
  using T = typename C::T;
  constexpr int V = C::V, L = C::L, M = C::M;
  if (tid >= C::kThreadsA) return;
  const int row = tid / M, t = tid - row * M;
  const long long pair = pair_base + row;
  CVec<V> x[L];
  if (pair < A.n_pairs) {
    const T* pd; const T* pa; const T* pd2 = nullptr; const T* pb;
    bool has_b = true;
    if (A.fused) {
      pd = reinterpret_cast<const T*>(A.donor) + pair * A.ld_donor;
      pa = reinterpret_cast<const T*>(A.dst_a) + pair * A.ld_a;
      pb = reinterpret_cast<const T*>(A.dst_b) + pair * A.ld_b;
      pd2 = pd;
    } else {
      const long long r0 = 2 * pair, r1 = 2 * pair + 1;
      pd = reinterpret_cast<const T*>(A.donor) + r0 * A.ld_donor;
      pa = reinterpret_cast<const T*>(A.dst_a) + r0 * A.ld_a;
      has_b = r1 < A.rows;
      pd2 = reinterpret_cast<const T*>(A.donor) + (has_b ? r1 : r0) * A.ld_donor;
      pb = reinterpret_cast<const T*>(A.dst_a) + (has_b ? r1 : r0) * A.ld_a;
    }
#pragma unroll
    for (int r = 0; r < L; ++r) {
      const int col = V * (t + M * r);
      float dn[V], a[V], b[V], dn2[V];
      Vld<T, V>::ld(pd + col, dn);
      Vld<T, V>::ld(pa + col, a);
      Vld<T, V>::ld(pb + col, b);
      if (A.fused) {
#pragma unroll
        for (int e = 0; e < V; ++e) dn2[e] = dn[e];
      } else {
        Vld<T, V>::ld(pd2 + col, dn2);
      }
#pragma unroll
      for (int e = 0; e < V; ++e) {
        x[r].re.set(e, dn[e] - a[e]);
        x[r].im.set(e, has_b ? dn2[e] - b[e] : 0.0f);
      }
    }
  } else {
#pragma unroll
    for (int r = 0; r < L; ++r)
#pragma unroll
      for (int e = 0; e < V; ++e) { x[r].re.set(e, 0.0f); x[r].im.set(e, 0.0f); }
  }
  fftreg::dft_inplace<L, false>(x);
#pragma unroll
  for (int k1 = 0; k1 < L; ++k1) {
    if (k1 > 0) {
      const float2 w = *reinterpret_cast<const float2*>(sm_tw1 + 2 * (k1 * M + t));
      x[k1] = fftreg::cmul(x[k1], w.x, w.y);
    }
    unit_store<C>(sm_exch + (row * L + k1) * C::kUnitFloats, t, x[k1]);
  }
}

// ---- phase B: M-point DFT, radix-V combine, filter, and back ---------------------------------------------
template <typename C>
VF_HD void phase_b(int tid, float* sm_exch, const float* sm_rec) {
  constexpr int V = C::V, L = C::L, M = C::M;
  if (tid >= C::kThreadsB) return;
  const int k1 = tid % L;
  float* unit = sm_exch + tid * C::kUnitFloats;      // tid == row*L + k1
  CVec<V> y[M];
#pragma unroll
  for (int t = 0; t < M; ++t) unit_load<C>(unit, t, y[t]);
  fftreg::dft_inplace<M, false>(y);
#pragma unroll
  for (int k2 = 0; k2 < M; ++k2) {
    const int kp = k1 + L * k2;
    const float* rec = sm_rec + kp * C::kRecFloats;
    if constexpr (V == 2) {
      float q[4];
      load4<V>(rec, q);                                // h0, h1, w.c, w.s
      const float z0r = y[k2].re.p[0].x, z0i = y[k2].im.p[0].x;
      const float y1r = y[k2].re.p[0].y, y1i = y[k2].im.p[0].y;
      const float z1r = y1r * q[2] - y1i * q[3], z1i = y1r * q[3] + y1i * q[2];
      const float x0r = (z0r + z1r) * q[0], x0i = (z0i + z1i) * q[0];
      const float x1r = (z0r - z1r) * q[1], x1i = (z0i - z1i) * q[1];
      const float u0r = x0r + x1r, u0i = x0i + x1i;
      const float u1r = x0r - x1r, u1i = x0i - x1i;
      // * conj(w)
      y[k2].re.p[0] = make_float2(u0r, u1r * q[2] + u1i * q[3]);
      y[k2].im.p[0] = make_float2(u0i, u1i * q[2] - u1r * q[3]);
    } else {
      float hq[4], w12[4], w3[4];
      load4<V>(rec, hq);
      load4<V>(rec + 4, w12);
      load4<V>(rec + 8, w3);
      const float y0r = y[k2].re.p[0].x, y0i = y[k2].im.p[0].x;
      const float y1r = y[k2].re.p[0].y, y1i = y[k2].im.p[0].y;
      const float y2r = y[k2].re.p[1].x, y2i = y[k2].im.p[1].x;
      const float y3r = y[k2].re.p[1].y, y3i = y[k2].im.p[1].y;
      const float z1r = y1r * w12[0] - y1i * w12[1], z1i = y1r * w12[1] + y1i * w12[0];
      const float z2r = y2r * w12[2] - y2i * w12[3], z2i = y2r * w12[3] + y2i * w12[2];
      const float z3r = y3r * w3[0] - y3i * w3[1], z3i = y3r * w3[1] + y3i * w3[0];
      // forward 4-point over the lanes
      float ar = y0r + z2r, ai = y0i + z2i, br = y0r - z2r, bi = y0i - z2i;
      float cr = z1r + z3r, ci = z1i + z3i, dr = z1i - z3i, di = -(z1r - z3r);   // (z1 - z3) * (-i)
      const float x0r = (ar + cr) * hq[0], x0i = (ai + ci) * hq[0];
      const float x1r = (br + dr) * hq[1], x1i = (bi + di) * hq[1];
      const float x2r = (ar - cr) * hq[2], x2i = (ai - ci) * hq[2];
      const float x3r = (br - dr) * hq[3], x3i = (bi - di) * hq[3];
      // inverse 4-point
      ar = x0r + x2r; ai = x0i + x2i; br = x0r - x2r; bi = x0i - x2i;
      cr = x1r + x3r; ci = x1i + x3i; dr = -(x1i - x3i); di = x1r - x3r;          // (x1 - x3) * (+i)
      const float u0r = ar + cr, u0i = ai + ci;
      const float u1r = br + dr, u1i = bi + di;
      const float u2r = ar - cr, u2i = ai - ci;
      const float u3r = br - dr, u3i = bi - di;
      // * conj(w_e)
      y[k2].re.p[0] = make_float2(u0r, u1r * w12[0] + u1i * w12[1]);
      y[k2].im.p[0] = make_float2(u0i, u1i * w12[0] - u1r * w12[1]);
      y[k2].re.p[1] = make_float2(u2r * w12[2] + u2i * w12[3], u3r * w3[0] + u3i * w3[1]);
      y[k2].im.p[1] = make_float2(u2i * w12[2] - u2r * w12[3], u3i * w3[0] - u3r * w3[1]);
    }
  }
  fftreg::dft_inplace<M, true>(y);
#pragma unroll
  for (int t = 0; t < M; ++t) unit_store<C>(unit, t, y[t]);
}

// ---- phase C: conj twiddle, inverse L-point DFT, add to dst, store ------------------------------------------
template <typename C>
VF_HD void phase_c(const Args& A, long long pair_base, int tid, const float* sm_exch, const float* sm_tw1) {
  using T = typename C::T;
  constexpr int V = C::V, L = C::L, M = C::M;
  if (tid >= C::kThreadsA) return;
  const int row = tid / M, t = tid - row * M;
  const long long pair = pair_base + row;
  if (pair >= A.n_pairs) return;
  const T* pa; T* oa; const T* pb; T* ob;
  bool has_b = true;
  if (A.fused) {
    pa = reinterpret_cast<const T*>(A.dst_a) + pair * A.ld_a;
    oa = reinterpret_cast<T*>(A.out_a) + pair * A.ld_out_a;
    pb = reinterpret_cast<const T*>(A.dst_b) + pair * A.ld_b;
    ob = reinterpret_cast<T*>(A.out_b) + pair * A.ld_out_b;
  } else {
    const long long r0 = 2 * pair, r1 = 2 * pair + 1;
    has_b = r1 < A.rows;
    pa = reinterpret_cast<const T*>(A.dst_a) + r0 * A.ld_a;
    oa = reinterpret_cast<T*>(A.out_a) + r0 * A.ld_out_a;
    pb = reinterpret_cast<const T*>(A.dst_a) + (has_b ? r1 : r0) * A.ld_a;
    ob = reinterpret_cast<T*>(A.out_a) + (has_b ? r1 : r0) * A.ld_out_a;
  }
  // dst is read again (an L2 hit: phase A touched it a few microseconds ago) instead of being carried in
  // registers through phase B; issue those loads FIRST so that their latency hides behind the inverse DFT.
  typename Vld<T, V>::Raw ra[L], rb[L];
#pragma unroll
  for (int r = 0; r < L; ++r) {
    const int col = V * (t + M * r);
    ra[r] = Vld<T, V>::ld_raw(pa + col);
    rb[r] = Vld<T, V>::ld_raw(pb + col);
  }
  CVec<V> x[L];
#pragma unroll
  for (int k1 = 0; k1 < L; ++k1) {
    unit_load<C>(sm_exch + (row * L + k1) * C::kUnitFloats, t, x[k1]);
    if (k1 > 0) {
      const float2 w = *reinterpret_cast<const float2*>(sm_tw1 + 2 * (k1 * M + t));
      x[k1] = fftreg::cmul(x[k1], w.x, -w.y);
    }
  }
  fftreg::dft_inplace<L, true>(x);
#pragma unroll
  for (int r = 0; r < L; ++r) {
    const int col = V * (t + M * r);
    float a[V], b[V];
    Vld<T, V>::unpack(ra[r], a);
    Vld<T, V>::unpack(rb[r], b);
#pragma unroll
    for (int e = 0; e < V; ++e) a[e] += x[r].re.get(e);
    Vld<T, V>::st(oa + col, a);
    if (has_b) {
#pragma unroll
      for (int e = 0; e < V; ++e) b[e] += x[r].im.get(e);
      Vld<T, V>::st(ob + col, b);
    }
  }
}

}  // namespace fsaifast
}  // namespace vf
