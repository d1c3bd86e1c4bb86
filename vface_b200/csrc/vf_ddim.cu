// Fused classifier-free-guidance + DDIM update (reverse step) and forward-DDIM (inversion) step.
//
// HBM-bound elementwise kernels: algorithmic traffic per latent element is 3 reads + 2 writes
// (x, e_u, e_c -> x_prev, pred_x0; + noise when sigma > 0).  16-byte vector accesses, grid sized
// to a multiple of the SM count; at B <= 256 frames the launch is latency-bound (459 KB/frame).
//
// Arithmetic follows the reference's fp32 op order without FMA contraction
// (ldm/models/diffusion/ddim_w_inv.py:666, :686, :696-700) so x_prev is bit-identical to the
// PyTorch eager result for fp32 epsilon inputs.
#include "vf_common.cuh"

#include <cmath>

namespace vf {

struct DdimScalars {
  float s1m_at;       // sqrt(1 - a_t)        (table entry)
  float sqrt_at;      // sqrt(a_t)
  float sqrt_aprev;   // sqrt(a_prev)
  float dir_coef;     // sqrt(1 - a_prev - sigma^2)
  float sigma;
  float cfg;
};

__device__ __forceinline__ void ddim_one(float x, float eu, float ec, float nz, const DdimScalars& s,
                                         float& x_prev, float& pred) {
  float e = __fadd_rn(eu, __fmul_rn(s.cfg, __fsub_rn(ec, eu)));
  pred = __fdiv_rn(__fsub_rn(x, __fmul_rn(s.s1m_at, e)), s.sqrt_at);
  float dir = __fmul_rn(s.dir_coef, e);
  float xp = __fadd_rn(__fmul_rn(s.sqrt_aprev, pred), dir);
  x_prev = __fadd_rn(xp, __fmul_rn(s.sigma, nz));
}

template <typename TE> struct Vec4;
template <> struct Vec4<float> {
  static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <> struct Vec4<__nv_bfloat16> {
  static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    v[0] = bf16lo(t.x); v[1] = bf16hi(t.x); v[2] = bf16lo(t.y); v[3] = bf16hi(t.y);
  }
};

template <typename TE, bool kNoise>
__global__ void __launch_bounds__(256)
ddim_cfg_kernel(const float* __restrict__ x, const TE* __restrict__ eu, const TE* __restrict__ ec,
                const float* __restrict__ noise, float* __restrict__ x_prev, float* __restrict__ pred_x0,
                long long n4, DdimScalars s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float xv[4], u[4], c[4], nz[4] = {0.f, 0.f, 0.f, 0.f}, xp[4], pr[4];
    Vec4<float>::ld(x + 4 * i, xv);
    Vec4<TE>::ld(eu + 4 * i, u);
    Vec4<TE>::ld(ec + 4 * i, c);
    if (kNoise) Vec4<float>::ld(noise + 4 * i, nz);
#pragma unroll
    for (int j = 0; j < 4; ++j) ddim_one(xv[j], u[j], c[j], nz[j], s, xp[j], pr[j]);
    *reinterpret_cast<float4*>(x_prev + 4 * i) = make_float4(xp[0], xp[1], xp[2], xp[3]);
    *reinterpret_cast<float4*>(pred_x0 + 4 * i) = make_float4(pr[0], pr[1], pr[2], pr[3]);
  }
}

struct InvScalars {
  float s1m_cur;     // sqrt(1 - a_cur)
  float sqrt_next;   // sqrt(a_next)
  float sqrt_cur;    // sqrt(a_cur)
  float s1m_next;    // sqrt(1 - a_next)
  float cfg;
};

template <typename TE, bool kCfg>
__global__ void __launch_bounds__(256)
ddim_invert_kernel(const float* __restrict__ x, const TE* __restrict__ eu, const TE* __restrict__ ec,
                   float* __restrict__ x_next, long long n4, InvScalars s) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    float xv[4], u[4] = {0.f, 0.f, 0.f, 0.f}, c[4], o[4];
    Vec4<float>::ld(x + 4 * i, xv);
    if (kCfg) Vec4<TE>::ld(eu + 4 * i, u);
    Vec4<TE>::ld(ec + 4 * i, c);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float e = kCfg ? __fadd_rn(u[j], __fmul_rn(s.cfg, __fsub_rn(c[j], u[j]))) : c[j];
      // ((x - sqrt(1-a_cur) e) * sqrt(a_next)) / sqrt(a_cur) + sqrt(1-a_next) e   (ddim_w_inv.py:449)
      float t = __fmul_rn(__fsub_rn(xv[j], __fmul_rn(s.s1m_cur, e)), s.sqrt_next);
      o[j] = __fadd_rn(__fdiv_rn(t, s.sqrt_cur), __fmul_rn(s.s1m_next, e));
    }
    *reinterpret_cast<float4*>(x_next + 4 * i) = make_float4(o[0], o[1], o[2], o[3]);
  }
}

static int grid_for(long long n4) {
  long long blocks = (n4 + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace vf

extern "C" int vf_ddim_cfg_step(const void* x, const void* e_uncond, const void* e_cond,
                                void* x_prev, void* pred_x0,
                                float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at,
                                float cfg_scale, const void* noise, long long n, int dtype_e, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !e_uncond || !e_cond || !x_prev || !pred_x0) return fail("vf_ddim_cfg_step: null pointer");
  if (n <= 0 || (n & 3)) return fail("vf_ddim_cfg_step: n=%lld must be a positive multiple of 4", n);
  if (dtype_e != VF_F32 && dtype_e != VF_BF16) return fail("vf_ddim_cfg_step: bad dtype %d", dtype_e);
  if (sigma_t != 0.f && !noise) return fail("vf_ddim_cfg_step: sigma_t != 0 needs a noise tensor");
  if (!aligned16(x) || !aligned16(x_prev) || !aligned16(pred_x0) || (noise && !aligned16(noise)) ||
      (reinterpret_cast<uintptr_t>(e_uncond) & 7) || (reinterpret_cast<uintptr_t>(e_cond) & 7))
    return fail("vf_ddim_cfg_step: pointers must be 16-byte aligned");
  DdimScalars s;
  s.s1m_at = sqrt_one_minus_at;
  s.sqrt_at = sqrtf(a_t);
  s.sqrt_aprev = sqrtf(a_prev);
  float t = 1.0f - a_prev;
  float sg2 = sigma_t * sigma_t;
  t = t - sg2;
  s.dir_coef = sqrtf(t);
  s.sigma = sigma_t;
  s.cfg = cfg_scale;
  long long n4 = n / 4;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n4);
  const float* xf = (const float*)x;
  const float* nz = (const float*)noise;
  float* xp = (float*)x_prev;
  float* pr = (float*)pred_x0;
  if (dtype_e == VF_F32) {
    if (noise) ddim_cfg_kernel<float, true><<<grid, 256, 0, st>>>(xf, (const float*)e_uncond, (const float*)e_cond, nz, xp, pr, n4, s);
    else       ddim_cfg_kernel<float, false><<<grid, 256, 0, st>>>(xf, (const float*)e_uncond, (const float*)e_cond, nz, xp, pr, n4, s);
  } else {
    if (noise) ddim_cfg_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(xf, (const __nv_bfloat16*)e_uncond, (const __nv_bfloat16*)e_cond, nz, xp, pr, n4, s);
    else       ddim_cfg_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(xf, (const __nv_bfloat16*)e_uncond, (const __nv_bfloat16*)e_cond, nz, xp, pr, n4, s);
  }
  return check_cuda(cudaGetLastError(), "ddim_cfg_kernel launch");
}

extern "C" int vf_ddim_invert_step(const void* x, const void* e_uncond, const void* e_cond, void* x_next,
                                   float a_cur, float a_next, float cfg_scale,
                                   long long n, int dtype_e, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!x || !e_cond || !x_next) return fail("vf_ddim_invert_step: null pointer");
  if (n <= 0 || (n & 3)) return fail("vf_ddim_invert_step: n=%lld must be a positive multiple of 4", n);
  if (dtype_e != VF_F32 && dtype_e != VF_BF16) return fail("vf_ddim_invert_step: bad dtype %d", dtype_e);
  if (!aligned16(x) || !aligned16(x_next)) return fail("vf_ddim_invert_step: pointers must be 16-byte aligned");
  // eps is read with 16-byte (fp32) / 8-byte (bf16) vector loads
  const uintptr_t emask = dtype_e == VF_F32 ? 15 : 7;
  if ((reinterpret_cast<uintptr_t>(e_cond) & emask) || (e_uncond && (reinterpret_cast<uintptr_t>(e_uncond) & emask)))
    return fail("vf_ddim_invert_step: e_cond/e_uncond must be %d-byte aligned", (int)emask + 1);
  InvScalars s;
  s.s1m_cur = sqrtf(1.0f - a_cur);
  s.sqrt_next = sqrtf(a_next);
  s.sqrt_cur = sqrtf(a_cur);
  s.s1m_next = sqrtf(1.0f - a_next);
  s.cfg = cfg_scale;
  long long n4 = n / 4;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(n4);
  const float* xf = (const float*)x;
  float* xo = (float*)x_next;
  if (dtype_e == VF_F32) {
    if (e_uncond) ddim_invert_kernel<float, true><<<grid, 256, 0, st>>>(xf, (const float*)e_uncond, (const float*)e_cond, xo, n4, s);
    else          ddim_invert_kernel<float, false><<<grid, 256, 0, st>>>(xf, nullptr, (const float*)e_cond, xo, n4, s);
  } else {
    if (e_uncond) ddim_invert_kernel<__nv_bfloat16, true><<<grid, 256, 0, st>>>(xf, (const __nv_bfloat16*)e_uncond, (const __nv_bfloat16*)e_cond, xo, n4, s);
    else          ddim_invert_kernel<__nv_bfloat16, false><<<grid, 256, 0, st>>>(xf, nullptr, (const __nv_bfloat16*)e_cond, xo, n4, s);
  }
  return check_cuda(cudaGetLastError(), "ddim_invert_kernel launch");
}
