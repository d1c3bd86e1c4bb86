// The UNet's OUTPUT convolution with an fp32 result (vf_conv3x3_out_f32).
//
//   eps = conv3x3(h, W) + b,  h: (n, hgt, wid, c) channels-last bf16, W: (4, c, 3, 3), eps: (n, 4, hgt, wid) fp32
//   (ldm/modules/diffusionmodules/openaimodel.py:835 `self.out[-1]`, evaluated at :907).
//
// Why a kernel of its own: the library convolution returns eps in bf16, and classifier-free guidance then forms
// e_u + s (e_c - e_u) (ddim_w_inv.py:666): the two branches' output roundings are independent, so the combination
// multiplies that last rounding by sqrt((s-1)^2 + s^2) = 3.6 at s = 3 -- measured as the largest single term of the
// bf16 path's per-step latent error (1.03e-2 -> below the 1e-2 bound with eps in fp32).  The convolution has four
// output channels (9 GFLOP per 96-sample step, 0.25 GB of input).
//
// Persistent CTAs (two per SM) stage the weights once, as ready-made B fragments of mma.sync.m16n8k16 (bf16 x bf16 ->
// fp32; the four output channels occupy half of N = 8), and walk 8x8 output tiles: the 10x10xc input halo of a tile
// (zero padded) lands in shared memory through 16-byte cp.async; warp = (16-pixel M-tile, half of the input channels);
// per tap and 16-channel block one ldmatrix.x4 (A: 16 pixels x 16 channels, rows at their tap-shifted halo addresses),
// one 8-byte B-fragment load and one MMA.  The first version did the 9 c 4 MACs per pixel with FMAs out of shared
// memory and was bound by shared-memory wavefronts at 0.60 ms per 96-sample step; the legacy tensor-core path needs
// only 1/9 of those reads.  (tcgen05 would be idle here: 9 GFLOP against 0.25 GB of input.)  Pixel stride c + 8
// elements keeps the eight 16-byte rows of every ldmatrix on distinct banks.
#include "vf_common.cuh"

namespace vf {

constexpr int kCoTile = 8;                 // 8x8 output pixels per CTA iteration
constexpr int kCoThreads = 256;            // 8 warps: 4 M-tiles (two output rows each) x 2 input-channel halves
constexpr int kCoOut = 4;

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__global__ void __launch_bounds__(kCoThreads)
conv_out_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w, const __nv_bfloat16* __restrict__ bias,
                float* __restrict__ out, int n_img, int hgt, int wid, int c) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int pstride = c + 8;                                   // elements per staged pixel (16-byte aligned, bank skew)
  const int kblocks = c / 16;
  __nv_bfloat16* s_in = reinterpret_cast<__nv_bfloat16*>(smem);                                   // [100][pstride]
  uint2* s_wf = reinterpret_cast<uint2*>(smem + (size_t)100 * pstride * sizeof(__nv_bfloat16));      // [9][kblocks][32]
  float* s_red = reinterpret_cast<float*>(s_wf + (size_t)9 * kblocks * 32);                          // [4][32][4]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int tiles_x = (wid + kCoTile - 1) / kCoTile, tiles_y = (hgt + kCoTile - 1) / kCoTile;
  const int n_tiles = n_img * tiles_y * tiles_x;

  // ---- B fragments of W[co][ci][ky][kx] (contiguous OIHW): lane (g = lane / 4, t = lane % 4) of block (tap, kb) holds
  //      b0 = W[g][16 kb + 2t, +1][tap], b1 = W[g][16 kb + 2t + 8, +9][tap]; output columns g >= 4 are zero -------------
  for (int i = tid; i < 9 * kblocks * 32; i += kCoThreads) {
    const int l = i & 31, blk = i >> 5;
    const int tap = blk / kblocks, kb = blk - tap * kblocks;
    const int g = l >> 2, t = l & 3;
    uint2 v = make_uint2(0u, 0u);
    if (g < kCoOut) {
      const __nv_bfloat16* wg = w + (size_t)g * c * 9 + tap;
      const int ci = 16 * kb + 2 * t;
      const uint32_t e0 = __bfloat16_as_ushort(wg[(size_t)ci * 9]), e1 = __bfloat16_as_ushort(wg[(size_t)(ci + 1) * 9]);
      const uint32_t e2 = __bfloat16_as_ushort(wg[(size_t)(ci + 8) * 9]), e3 = __bfloat16_as_ushort(wg[(size_t)(ci + 9) * 9]);
      v.x = e0 | (e1 << 16);
      v.y = e2 | (e3 << 16);
    }
    s_wf[i] = v;
  }
  const int chunks = c / 8;
  const int mt = warp & 3, half = warp >> 2;                   // M-tile (output rows 2 mt, 2 mt + 1), input-channel half
  const int kb0 = half * (kblocks / 2), kb1 = kb0 + kblocks / 2;
  // ldmatrix row of this lane: matrix (lane / 8): rows 0-7 | 8-15 of the M-tile, channels +0 | +8
  const int lrow = (lane & 7) + ((lane >> 3) & 1) * 8, lk = (lane >> 4) * 8;
  const int l_oy = 2 * mt + (lrow >> 3), l_ox = lrow & 7;
  const uint32_t s_in_addr = static_cast<uint32_t>(__cvta_generic_to_shared(s_in));
  const int g = lane >> 2, t = lane & 3;
  const float b_lo = (bias && t < 2) ? __bfloat162float(bias[2 * t]) : 0.f;
  const float b_hi = (bias && t < 2) ? __bfloat162float(bias[2 * t + 1]) : 0.f;

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int n = tile / (tiles_y * tiles_x);
    const int ty = ((tile / tiles_x) % tiles_y) * kCoTile, tx = (tile % tiles_x) * kCoTile;
    __syncthreads();                                           // previous tile's halo and partial sums fully consumed
    // ---- stage the input halo: 16-byte cp.async per chunk, out-of-image pixels zero-filled (source size 0) -------
    const __nv_bfloat16* xn = x + (size_t)n * hgt * wid * c;
    for (int i = tid; i < 100 * chunks; i += kCoThreads) {
      const int pix = i / chunks, ch = (i - pix * chunks) * 8;
      const int py = ty + pix / 10 - 1, px = tx + pix % 10 - 1;
      const bool inside = py >= 0 && py < hgt && px >= 0 && px < wid;
      const __nv_bfloat16* src = inside ? xn + ((size_t)py * wid + px) * c + ch : xn;
      const uint32_t dst = s_in_addr + (uint32_t)(((size_t)pix * pstride + ch) * sizeof(__nv_bfloat16));
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" :: "r"(dst), "l"(src), "r"(inside ? 16 : 0) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int tap = 0; tap < 9; ++tap) {
      const int pix = (l_oy + tap / 3) * 10 + (l_ox + tap % 3);
      const uint32_t a_base = s_in_addr + (uint32_t)(((size_t)pix * pstride + lk) * sizeof(__nv_bfloat16));
      const uint2* wf = s_wf + (size_t)tap * kblocks * 32 + lane;
      for (int kb = kb0; kb < kb1; ++kb) {
        uint32_t a0, a1, a2, a3;
        ldmatrix_x4(a_base + (uint32_t)kb * 32u, a0, a1, a2, a3);
        const uint2 bf = wf[kb * 32];
        mma_16816(acc, a0, a1, a2, a3, bf.x, bf.y);
      }
    }
    // ---- combine the two channel halves (fixed order: bit-reproducible), add the bias, store NCHW fp32 -----------
    if (half == 1) *reinterpret_cast<float4*>(s_red + ((size_t)mt * 32 + lane) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    __syncthreads();
    if (half == 0 && t < 2) {
      const float4 o = *reinterpret_cast<const float4*>(s_red + ((size_t)mt * 32 + lane) * 4);
      const size_t plane = (size_t)hgt * wid;
      const int xx = tx + g;
      float* on = out + (size_t)n * kCoOut * plane;
      if (xx < wid) {
        const int y0 = ty + 2 * mt, y1 = y0 + 1;
        if (y0 < hgt) {
          on[(size_t)(2 * t) * plane + (size_t)y0 * wid + xx] = acc[0] + o.x + b_lo;
          on[(size_t)(2 * t + 1) * plane + (size_t)y0 * wid + xx] = acc[1] + o.y + b_hi;
        }
        if (y1 < hgt) {
          on[(size_t)(2 * t) * plane + (size_t)y1 * wid + xx] = acc[2] + o.z + b_lo;
          on[(size_t)(2 * t + 1) * plane + (size_t)y1 * wid + xx] = acc[3] + o.w + b_hi;
        }
      }
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
  const size_t smem = (size_t)100 * (c + 8) * sizeof(__nv_bfloat16) + (size_t)9 * (c / 16) * 32 * sizeof(uint2) + 4 * 32 * 4 * sizeof(float);
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
