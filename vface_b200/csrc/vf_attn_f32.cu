// fp32 reference-precision attention core (dtype VF_F32 of vf_attn_fwd): streaming softmax on CUDA
// cores, fp32 storage and arithmetic.  This is the precision-parity path (per-step latents against
// the fp32 reference, SURVEY.md section 4 level S1); the throughput path is the tcgen05 kernel in
// vf_attn_tc.cu.  Replaces ldm/models/pnp_utils.py:270-286 without materialising the N x N matrix.
//
// Mapping: a CTA of 128 threads owns 32 query rows of one (batch, head); 4 lanes share a row, each
// holding an interleaved quarter of the head dimension of q and of the output accumulator.  Key/value
// tiles of 32 rows are staged in shared memory; scores are reduced across the 4 lanes by shuffles and
// the softmax is updated online once per 8 keys.
#include "vf_attn.cuh"

#include <cfloat>

namespace vf {

constexpr int kF32Threads = 128;
constexpr int kF32Rows = 32;     // query rows per CTA
constexpr int kF32Keys = 32;     // keys per shared-memory tile
constexpr int kF32Chunk = 8;     // keys per online-softmax update


template <int DPT>   // head dims per thread (d <= 4*DPT)
__global__ void __launch_bounds__(kF32Threads)
attn_f32_kernel(const AttnF32Params P) {
  extern __shared__ __align__(16) float smem_f[];
  const int d = P.d;
  float* sk = smem_f;                 // kF32Keys * d
  float* sv = smem_f + kF32Keys * d;  // kF32Keys * d

  const int bh = blockIdx.y;
  const int b = bh / P.heads, h = bh - b * P.heads;
  const int row = blockIdx.x * kF32Rows + threadIdx.x / 4;
  const int sl = threadIdx.x & 3;
  const bool row_ok = row < P.n_q;
  const int nd = (d - sl + 3) / 4;    // dims 4i+sl < d

  float qr[DPT], acc[DPT];
#pragma unroll
  for (int i = 0; i < DPT; ++i) {
    acc[i] = 0.0f;
    qr[i] = (row_ok && i < nd) ? P.q[((long long)b * P.n_q + row) * P.ld_q + h * d + 4 * i + sl] : 0.0f;
  }
  float m = -FLT_MAX, l = 0.0f;

  for (int seg = 0; seg < 2; ++seg) {
    const float* kp = seg == 0 ? P.k : P.k2;
    const float* vp = seg == 0 ? P.v : P.v2;
    const int nk = seg == 0 ? P.n_kv : P.n_kv2;
    const long long ldk = seg == 0 ? P.ld_k : P.ld_k2;
    const long long ldv = seg == 0 ? P.ld_v : P.ld_v2;
    if (kp == nullptr || nk <= 0) continue;
    for (int k0 = 0; k0 < nk; k0 += kF32Keys) {
      const int kt = min(kF32Keys, nk - k0);
      __syncthreads();
      for (int idx = threadIdx.x; idx < kt * d; idx += kF32Threads) {
        const int r = idx / d, c = idx - r * d;
        sk[idx] = kp[((long long)b * nk + k0 + r) * ldk + h * d + c];
        sv[idx] = vp[((long long)b * nk + k0 + r) * ldv + h * d + c];
      }
      __syncthreads();
      for (int c0 = 0; c0 < kt; c0 += kF32Chunk) {
        float s[kF32Chunk];
        float cmax = -FLT_MAX;
#pragma unroll
        for (int j = 0; j < kF32Chunk; ++j) {
          float part = 0.0f;
          if (c0 + j < kt) {
            const float* kr = sk + (c0 + j) * d + sl;
#pragma unroll
            for (int i = 0; i < DPT; ++i)
              if (i < nd) part = fmaf(qr[i], kr[4 * i], part);
          }
          part += __shfl_xor_sync(0xffffffffu, part, 1);
          part += __shfl_xor_sync(0xffffffffu, part, 2);
          s[j] = (c0 + j < kt) ? part * P.scale : -FLT_MAX;
          cmax = fmaxf(cmax, s[j]);
        }
        const float m_new = fmaxf(m, cmax);
        const float corr = expf(m - m_new);
        l *= corr;
#pragma unroll
        for (int i = 0; i < DPT; ++i) acc[i] *= corr;
#pragma unroll
        for (int j = 0; j < kF32Chunk; ++j) {
          if (c0 + j < kt) {
            const float pj = expf(s[j] - m_new);
            l += pj;
            const float* vr = sv + (c0 + j) * d + sl;
#pragma unroll
            for (int i = 0; i < DPT; ++i)
              if (i < nd) acc[i] = fmaf(pj, vr[4 * i], acc[i]);
          }
        }
        m = m_new;
      }
    }
  }
  if (row_ok) {
    const float inv = 1.0f / l;
#pragma unroll
    for (int i = 0; i < DPT; ++i)
      if (i < nd) P.o[((long long)b * P.n_q + row) * P.ld_o + h * d + 4 * i + sl] = acc[i] * inv;
  }
}

int launch_attn_f32(const AttnF32Params& P, int batch, cudaStream_t st) {
  const int d = P.d;
  dim3 grid((P.n_q + kF32Rows - 1) / kF32Rows, batch * P.heads);
  const size_t smem = 2 * (size_t)kF32Keys * d * sizeof(float);
  if (d <= 64) {
    attn_f32_kernel<16><<<grid, kF32Threads, smem, st>>>(P);
  } else if (d <= 160) {
    attn_f32_kernel<40><<<grid, kF32Threads, smem, st>>>(P);
  } else {
    static bool attr = false;
    if (!attr) { VF_CUDA_TRY(cudaFuncSetAttribute(attn_f32_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024)); attr = true; }
    attn_f32_kernel<64><<<grid, kF32Threads, smem, st>>>(P);
  }
  return check_cuda(cudaGetLastError(), "attn_f32_kernel launch");
}

}  // namespace vf
