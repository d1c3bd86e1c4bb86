// Flow-Guided Attention Temporal Smoothening: bilinear border-padded gather of the previous
// frame's attention features along a precomputed flow, blended with the current frame.
//
// Layout: native token layout (frames, h*w, c) -- one pixel's c channels are contiguous, so each of
// the four bilinear taps is a contiguous row segment and every access is a coalesced 16-byte vector.
// The reference permutes to (b, c, h, w), calls grid_sample per frame pair in a Python loop and
// permutes back (ldm/models/pnp_utils.py:206-218, scripts/temporal_flow.py:40-53, :222-237); here one
// launch covers all frames with no layout round trip.
//
// HBM-bound: algorithmic bytes per frame = 3*N*C*e + 8*N (read x[i+1], gather x[i] ~ once through
// L2, write out, read flow).  The tap index chain is fp32 in exactly the reference's op order with
// no FMA contraction and no fast-math, so floor indices are bit-exact with torch grid_sample.
#include "vf_common.cuh"

namespace vf {

struct Taps {
  int x0, y0;
  float w_nw, w_ne, w_sw, w_se;
  bool in_x, in_y;
};

__device__ __forceinline__ float unnormalized_coord(float p, float f, float size_m1_safe, float size_m1) {
  // temporal_flow.py:45-49:  v = 2.0 * (p + f) / max(size-1, 1) - 1.0
  float g = __fadd_rn(p, f);
  float v = __fsub_rn(__fdiv_rn(__fmul_rn(2.0f, g), size_m1_safe), 1.0f);
  // ATen grid_sampler_unnormalize(align_corners=True): ((v + 1) / 2) * (size - 1)
  float i = __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.0f), 2.0f), size_m1);
  // padding_mode='border': clip_coordinates
  return fminf(size_m1, fmaxf(i, 0.0f));
}

__device__ __forceinline__ Taps make_taps(int px, int py, float fx, float fy, int h, int w) {
  Taps t;
  float wm1 = (float)(w - 1), hm1 = (float)(h - 1);
  float ix = unnormalized_coord((float)px, fx, fmaxf(wm1, 1.0f), wm1);
  float iy = unnormalized_coord((float)py, fy, fmaxf(hm1, 1.0f), hm1);
  float x0f = floorf(ix), y0f = floorf(iy);
  float x1f = __fadd_rn(x0f, 1.0f), y1f = __fadd_rn(y0f, 1.0f);
  t.x0 = (int)x0f;
  t.y0 = (int)y0f;
  t.w_nw = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(y1f, iy));
  t.w_ne = __fmul_rn(__fsub_rn(ix, x0f), __fsub_rn(y1f, iy));
  t.w_sw = __fmul_rn(__fsub_rn(x1f, ix), __fsub_rn(iy, y0f));
  t.w_se = __fmul_rn(__fsub_rn(ix, x0f), __fsub_rn(iy, y0f));
  t.in_x = (t.x0 + 1) <= (w - 1);
  t.in_y = (t.y0 + 1) <= (h - 1);
  return t;
}

template <typename T> struct Chunk;   // one 16-byte chunk as floats
template <> struct Chunk<float> {
  static constexpr int kElems = 4;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&v)[4]) {
    v[0] = __uint_as_float(u.x); v[1] = __uint_as_float(u.y);
    v[2] = __uint_as_float(u.z); v[3] = __uint_as_float(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[4]) {
    return make_uint4(__float_as_uint(v[0]), __float_as_uint(v[1]), __float_as_uint(v[2]), __float_as_uint(v[3]));
  }
};
template <> struct Chunk<__nv_bfloat16> {
  static constexpr int kElems = 8;
  static __device__ __forceinline__ void unpack(const uint4& u, float (&v)[8]) {
    v[0] = bf16lo(u.x); v[1] = bf16hi(u.x); v[2] = bf16lo(u.y); v[3] = bf16hi(u.y);
    v[4] = bf16lo(u.z); v[5] = bf16hi(u.z); v[6] = bf16lo(u.w); v[7] = bf16hi(u.w);
  }
  static __device__ __forceinline__ uint4 pack(const float (&v)[8]) {
    return make_uint4(pack_bf16(v[0], v[1]), pack_bf16(v[2], v[3]), pack_bf16(v[4], v[5]), pack_bf16(v[6], v[7]));
  }
};

// ---- blend arithmetic on one 16-byte chunk ----------------------------------------------------------
// Values follow grid_sample's tap order (nw, ne, sw, se) and temporal_flow.py:234 with fused
// multiply-adds (packed FFMA2): within 1 ulp of the reference's separate mul/add chain; only the tap
// INDEX chain above has to be (and is) bit-exact.
struct Wts {
  float nw, ne, sw, se;
};

template <typename T>
__device__ __forceinline__ uint4 blend_chunk(const uint4& cu, const uint4& u_nw, const uint4& u_ne, const uint4& u_sw,
                                              const uint4& u_se, const Wts& wt, float alpha, float one_minus_alpha) {
  constexpr int E = Chunk<T>::kElems;
  float c[E], a[E], b[E], s[E], d[E], o[E];
  Chunk<T>::unpack(cu, c);
  Chunk<T>::unpack(u_nw, a);
  Chunk<T>::unpack(u_ne, b);
  Chunk<T>::unpack(u_sw, s);
  Chunk<T>::unpack(u_se, d);
#pragma unroll
  for (int j = 0; j < E; j += 2) {
    float2 acc = __fmul2_rn(make_float2(a[j], a[j + 1]), make_float2(wt.nw, wt.nw));
    acc = __ffma2_rn(make_float2(b[j], b[j + 1]), make_float2(wt.ne, wt.ne), acc);
    acc = __ffma2_rn(make_float2(s[j], s[j + 1]), make_float2(wt.sw, wt.sw), acc);
    acc = __ffma2_rn(make_float2(d[j], d[j + 1]), make_float2(wt.se, wt.se), acc);
    const float2 ac = __fmul2_rn(make_float2(c[j], c[j + 1]), make_float2(alpha, alpha));
    const float2 r = __ffma2_rn(acc, make_float2(one_minus_alpha, one_minus_alpha), ac);
    o[j] = r.x;
    o[j + 1] = r.y;
  }
  return Chunk<T>::pack(o);
}

// One CTA iteration = one 8x8 pixel tile of one frame.  The 64 pixels' taps (four source-pixel indices
// + four weights) are computed ONCE by 64 threads into shared memory -- the index chain costs ~100
// instructions (two IEEE divisions per axis) and used to be repeated by every 16-byte chunk of the
// pixel (40x at C=320 bf16), which made the kernel issue-bound.  Then 8 lanes per pixel stream the
// channel chunks: 128-byte segments per tap row, the 2-D tile keeps the gathered rows L1/L2-resident.
constexpr int kWarpThreads = 256;
constexpr int kTile = 8;
constexpr int kLanesPerPx = 8;

template <typename T>
__global__ void __launch_bounds__(kWarpThreads)
flow_warp_blend_kernel(const T* __restrict__ x, const T* __restrict__ halo, const float* __restrict__ flow,
                       T* __restrict__ out, int* __restrict__ taps_out,
                       int frames, int h, int w, int chunks_per_px,
                       long long ld_x, long long ld_halo, long long ld_out, float alpha, float one_minus_alpha) {
  constexpr int E = Chunk<T>::kElems;
  __shared__ int4 s_off[kTile * kTile];
  __shared__ float4 s_wt[kTile * kTile];
  const int npx = h * w;
  const int tiles_x = (w + kTile - 1) / kTile, tiles_y = (h + kTile - 1) / kTile;
  const int tiles_per_frame = tiles_x * tiles_y;
  const long long n_tiles = (long long)tiles_per_frame * frames;
  const bool has_halo = halo != nullptr;
  const int cl = threadIdx.x & (kLanesPerPx - 1);
  const int ps = threadIdx.x / kLanesPerPx;                  // 0..31

  for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int f = (int)(tile / tiles_per_frame);
    const int tr = (int)(tile - (long long)f * tiles_per_frame);
    const int ty0 = (tr / tiles_x) * kTile, tx0 = (tr % tiles_x) * kTile;
    const bool copy_only = (f == 0 && !has_halo);            // out[0] = x[0] (temporal_flow.py:229)
    const int fl = has_halo ? f : f - 1;
    if (!copy_only && threadIdx.x < kTile * kTile) {
      const int py = ty0 + (threadIdx.x >> 3), px = tx0 + (threadIdx.x & 7);
      if (py < h && px < w) {
        const int p = py * w + px;
        const float* fp = flow + (long long)fl * 2 * npx;
        const Taps t = make_taps(px, py, __ldg(fp + p), __ldg(fp + npx + p), h, w);
        if (taps_out != nullptr) {
          taps_out[((long long)fl * npx + p) * 2 + 0] = t.x0;
          taps_out[((long long)fl * npx + p) * 2 + 1] = t.y0;
        }
        const int x1 = t.in_x ? t.x0 + 1 : t.x0, y1 = t.in_y ? t.y0 + 1 : t.y0;
        s_off[threadIdx.x] = make_int4(t.y0 * w + t.x0, t.y0 * w + x1, y1 * w + t.x0, y1 * w + x1);
        s_wt[threadIdx.x] = make_float4(t.w_nw, t.in_x ? t.w_ne : 0.0f, t.in_y ? t.w_sw : 0.0f,
                                        (t.in_x && t.in_y) ? t.w_se : 0.0f);
      }
    }
    __syncthreads();
    const T* prev;
    long long ldp;
    if (f == 0) { prev = halo; ldp = ld_halo; }
    else        { prev = x + (long long)(f - 1) * npx * ld_x; ldp = ld_x; }
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int tp = ps + half * 32;
      const int py = ty0 + (tp >> 3), px = tx0 + (tp & 7);
      if (py >= h || px >= w) continue;
      const long long p = (long long)f * npx + py * w + px;
      const T* cur = x + p * ld_x;
      T* dst = out + p * ld_out;
      if (copy_only) {
        for (int ch = cl; ch < chunks_per_px; ch += kLanesPerPx) st_na_v4(dst + ch * E, ld_nc_v4(cur + ch * E));
        continue;
      }
      const int4 off = s_off[tp];
      const float4 wq = s_wt[tp];
      const Wts wt{wq.x, wq.y, wq.z, wq.w};
      const T* r_nw = prev + (long long)off.x * ldp;
      const T* r_ne = prev + (long long)off.y * ldp;
      const T* r_sw = prev + (long long)off.z * ldp;
      const T* r_se = prev + (long long)off.w * ldp;
#pragma unroll 2
      for (int ch = cl; ch < chunks_per_px; ch += kLanesPerPx) {
        const int e0 = ch * E;
        const uint4 cu = ld_nc_v4(cur + e0);
        // gathered rows are re-read by neighbouring pixels: default (cached) loads
        const uint4 u_nw = *reinterpret_cast<const uint4*>(r_nw + e0);
        const uint4 u_ne = *reinterpret_cast<const uint4*>(r_ne + e0);
        const uint4 u_sw = *reinterpret_cast<const uint4*>(r_sw + e0);
        const uint4 u_se = *reinterpret_cast<const uint4*>(r_se + e0);
        st_na_v4(dst + e0, blend_chunk<T>(cu, u_nw, u_ne, u_sw, u_se, wt, alpha, one_minus_alpha));
      }
    }
    __syncthreads();
  }
}

}  // namespace vf

extern "C" int vf_flow_warp_blend(const void* x, const void* prev_halo, const float* flow, void* out,
                                  int frames, int h, int w, int c,
                                  long long ld_x, long long ld_halo, long long ld_out,
                                  double alpha, int dtype, int* taps_out, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !out) return fail("vf_flow_warp_blend: null pointer");
  if (x == out) return fail("vf_flow_warp_blend: out must not alias x (frame i+1 reads the un-aligned frame i)");
  if (frames <= 0 || h <= 0 || w <= 0 || c <= 0) return fail("vf_flow_warp_blend: bad shape frames=%d h=%d w=%d c=%d", frames, h, w, c);
  if (dtype != VF_F32 && dtype != VF_BF16) return fail("vf_flow_warp_blend: bad dtype %d", dtype);
  const int esize = dtype == VF_F32 ? 4 : 2;
  const int epc = 16 / esize;
  if (c % epc) return fail("vf_flow_warp_blend: c=%d * %d bytes must be a multiple of 16", c, esize);
  if (ld_x < c || ld_out < c || (ld_x % epc) || (ld_out % epc)) return fail("vf_flow_warp_blend: bad row strides");
  if (prev_halo && (ld_halo < c || (ld_halo % epc))) return fail("vf_flow_warp_blend: bad halo stride");
  if ((frames > 1 || prev_halo) && !flow) return fail("vf_flow_warp_blend: flow is null");
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 15) ||
      (prev_halo && (reinterpret_cast<uintptr_t>(prev_halo) & 15)))
    return fail("vf_flow_warp_blend: pointers must be 16-byte aligned");
  const int cpp = c / epc;
  long long blocks = (long long)frames * ((h + vf::kTile - 1) / vf::kTile) * ((w + vf::kTile - 1) / vf::kTile);
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  const float a = (float)alpha;
  const float b = (float)(1.0 - alpha);   // python: (1 - alpha) in double, cast to fp32 by the tensor op
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VF_F32)
    flow_warp_blend_kernel<float><<<(int)blocks, kWarpThreads, 0, st>>>((const float*)x, (const float*)prev_halo, flow, (float*)out,
                                                               taps_out, frames, h, w, cpp, ld_x, ld_halo, ld_out, a, b);
  else
    flow_warp_blend_kernel<__nv_bfloat16><<<(int)blocks, kWarpThreads, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)prev_halo, flow,
                                                                       (__nv_bfloat16*)out, taps_out, frames, h, w, cpp,
                                                                       ld_x, ld_halo, ld_out, a, b);
  return check_cuda(cudaGetLastError(), "flow_warp_blend_kernel launch");
}
