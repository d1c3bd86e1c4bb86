// The UNet's OUTPUT convolution with an fp32 result (vf_conv3x3_out_f32).
//
//   eps = conv3x3(h, W) + b,  h: (n, hgt, wid, c) channels-last bf16, W: (4, c, 3, 3), eps: (n, 4, hgt, wid) fp32
//   (ldm/modules/diffusionmodules/openaimodel.py:835 `self.out[-1]`, evaluated at :907).
//
// Why a kernel of its own: the library convolution returns eps in bf16, and classifier-free guidance then forms
// e_u + s (e_c - e_u) (ddim_w_inv.py:666): the two branches' output roundings are independent, so the combination
// multiplies that last rounding by sqrt((s-1)^2 + s^2) = 3.6 at s = 3 -- measured as the largest single term of the
// bf16 path's per-step latent error (1.03e-2 -> below the 1e-2 bound with eps in fp32).  The convolution has four
// output channels (9 GFLOP per 96-sample step, 0.25 GB of input): CUDA-core work, bound by shared-memory reads.
//
// Persistent CTAs (two per SM) stage the weights once (fp32, [tap][slice][ci] x 4 outputs) and walk 8x8 output tiles:
// the 10x10xc input halo of a tile (zero padded) goes to shared memory; thread = (pixel, quarter of the input
// channels), four fp32 accumulators (the four output channels), quarter sums combined with two shuffles.  Pixel stride c+8 elements and slice stride +1 weight row keep the 16-byte
// shared-memory reads of a quarter-warp on distinct banks.
#include "vf_common.cuh"

namespace vf {

constexpr int kCoTile = 8;                 // 8x8 output pixels per CTA
constexpr int kCoThreads = 256;            // 64 pixels x 4 input-channel slices
constexpr int kCoOut = 4;

__global__ void __launch_bounds__(kCoThreads)
conv_out_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
                float* __restrict__ out, int n_img, int hgt, int wid, int c) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int pstride = c + 8;                                   // elements per staged pixel (16-byte aligned, bank skew)
  const int slice = c / 4;                                     // input channels per slice (multiple of 8)
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(smem);                                  // [100][pstride]
  float4* s_w = reinterpret_cast<float4*>(smem + (size_t)100 * pstride * sizeof(__nv_bfloat16));    // [9][4][slice + 1]
  const int tid = threadIdx.x;
  const int tiles_x = (wid + kCoTile - 1) / kCoTile, tiles_y = (hgt + kCoTile - 1) / kCoTile;
  const int n_tiles = n_img * tiles_y * tiles_x;

  // ---- stage the weights: W[co][ci][ky][kx] (contiguous OIHW) -> s_w[tap][s][ci_local] = (co 0..3) ----------
  for (int i = tid; i < 9 * c; i += kCoThreads) {
    const int tap = i / c, ci = i - tap * c;
    float4 v;
    v.x = __bfloat162float(w[((size_t)0 * c + ci) * 9 + tap]);
    v.y = __bfloat162float(w[((size_t)1 * c + ci) * 9 + tap]);
    v.z = __bfloat162float(w[((size_t)2 * c + ci) * 9 + tap]);
    v.w = __bfloat162float(w[((size_t)3 * c + ci) * 9 + tap]);
    const int s = ci / slice;
    s_w[(tap * 4 + s) * (slice + 1) + (ci - s * slice)] = v;
  }
  const int chunks = c / 8;
  const int p = tid >> 2, s = tid & 3;                         // pixel of the tile, input-channel slice
  const int oy = p >> 3, ox = p & 7;
  const float b0 = bias ? __bfloat162float(bias[0]) : 0.f, b1 = bias ? __bfloat162float(bias[1]) : 0.f;
  const float b2 = bias ? __bfloat162float(bias[2]) : 0.f, b3 = bias ? __bfloat162float(bias[3]) : 0.f;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
  const int n = tile / (tiles_y * tiles_x);
  const int ty = ((tile / tiles_x) % tiles_y) * kCoTile, tx = (tile % tiles_x) * kCoTile;
  __syncthreads();                                             // previous tile's halo fully consumed
  // ---- stage the input halo (zero outside the image) ---------------------------------------------------------
  const __nv_bfloat16* xn = x + (size_t)n * hgt * wid * c;
  for (int i = tid; i < 100 * chunks; i += kCoThreads) {
    const int pix = i / chunks, ch = (i - pix * chunks) * 8;
    const int py = ty + pix / 10 - 1, px = tx + pix % 10 - 1;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (py >= 0 && py < hgt && px >= 0 && px < wid) v = ld_nc_v4(xn + ((size_t)py * wid + px) * c + ch);
    *reinterpret_cast<uint4*>(s_in + (size_t)pix * pstride + ch) = v;
  }
  __syncthreads();

  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
  for (int tap = 0; tap < 9; ++tap) {
    const int pix = (oy + tap / 3) * 10 + (ox + tap % 3);
    const uint4* in = reinterpret_cast<const uint4*>(s_in + (size_t)pix * pstride + s * slice);
    const float4* wv = s_w + (tap * 4 + s) * (slice + 1);
    for (int k = 0; k < slice / 8; ++k) {
      const uint4 raw = in[k];
      const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float lo = bf16lo(r[q]), hi = bf16hi(r[q]);
        const float4 w0 = wv[k * 8 + 2 * q], w1 = wv[k * 8 + 2 * q + 1];
        a0 = fmaf(lo, w0.x, a0); a1 = fmaf(lo, w0.y, a1); a2 = fmaf(lo, w0.z, a2); a3 = fmaf(lo, w0.w, a3);
        a0 = fmaf(hi, w1.x, a0); a1 = fmaf(hi, w1.y, a1); a2 = fmaf(hi, w1.z, a2); a3 = fmaf(hi, w1.w, a3);
      }
    }
  }
  // combine the four slices (lanes 4p .. 4p+3), fixed order: bit-reproducible
#pragma unroll
  for (int off = 1; off <= 2; off <<= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, off);
    a1 += __shfl_xor_sync(0xffffffffu, a1, off);
    a2 += __shfl_xor_sync(0xffffffffu, a2, off);
    a3 += __shfl_xor_sync(0xffffffffu, a3, off);
  }
  const int y = ty + oy, xx = tx + ox;
  if (s == 0 && y < hgt && xx < wid) {
    const size_t plane = (size_t)hgt * wid;
    float* o = out + (size_t)n * kCoOut * plane + (size_t)y * wid + xx;
    o[0 * plane] = a0 + b0;
    o[1 * plane] = a1 + b1;
    o[2 * plane] = a2 + b2;
    o[3 * plane] = a3 + b3;
  }
  }   // tile loop
}

}  // namespace vf

extern "C" int vf_conv3x3_out_f32(const void* x, const void* weight, const void* bias, void* out, int n, int h, int w, int c,
                                  int c_out, int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (dtype != VF_BF16) return fail("vf_conv3x3_out_f32: bf16 activations only (the fp32 path keeps the library convolution), got dtype %d", dtype);
  if (c_out != kCoOut) return fail("vf_conv3x3_out_f32: c_out=%d, the UNet output convolution has %d channels", c_out, kCoOut);
  if (c % 32 || c <= 0) return fail("vf_conv3x3_out_f32: c=%d must be a positive multiple of 32", c);
  if (n <= 0 || h <= 0 || w <= 0) return fail("vf_conv3x3_out_f32: bad shape n=%d h=%d w=%d", n, h, w);
  if ((reinterpret_cast<uintptr_t>(x) & 15) || (reinterpret_cast<uintptr_t>(out) & 3))
    return fail("vf_conv3x3_out_f32: x must be 16-byte aligned");
  const size_t smem = (size_t)100 * (c + 8) * sizeof(__nv_bfloat16) + (size_t)9 * 4 * (c / 4 + 1) * sizeof(float4);
  if (smem > 200 * 1024) return fail("vf_conv3x3_out_f32: c=%d needs %zu bytes of shared memory", c, smem);
  static size_t attr = 0;
  if (smem > attr) {
    VF_CUDA_TRY(cudaFuncSetAttribute(conv_out_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = smem;
  }
  const long long tiles = (long long)n * ((h + kCoTile - 1) / kCoTile) * ((w + kCoTile - 1) / kCoTile);
  const int per_sm = smem <= 113 * 1024 ? 2 : 1;
  const int grid = (int)(tiles < (long long)per_sm * num_sms() ? tiles : (long long)per_sm * num_sms());
  conv_out_kernel<<<grid, kCoThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(weight),
      reinterpret_cast<const __nv_bfloat16*>(bias), reinterpret_cast<float*>(out), n, h, w, c);
  return check_cuda(cudaGetLastError(), "conv_out_kernel launch");
}
