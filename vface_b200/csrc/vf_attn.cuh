// Declarations shared by the attention dispatcher and its two kernels.
#pragma once
#include "vf_common.cuh"

namespace vf {

struct AttnF32Params {
  const float* q; const float* k; const float* v; float* o;
  const float* k2; const float* v2;
  int heads, n_q, n_kv, n_kv2, d;
  long long ld_q, ld_k, ld_v, ld_o, ld_k2, ld_v2;
  float scale;
};

int launch_attn_f32(const AttnF32Params& P, int batch, cudaStream_t st);
int launch_attn_tc(const void* q, const void* k, const void* v, void* o, int batch, int heads, int n_q, int n_kv,
                   int d, long long ld_q, long long ld_k, long long ld_v, long long ld_o, float scale,
                   const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2, cudaStream_t st);

int launch_attn_stream(const void* q, const void* k, const void* v, void* o, int batch, int heads, int n_q, int n_kv,
                       int d, long long ld_q, long long ld_k, long long ld_v, long long ld_o, float scale,
                       const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2, int emu,
                       cudaStream_t st);

// round-robin arrangement (vf_attn_pp.cu): one CTA per SM, three query tiles, softmax warps take turns on the XU; d_head <= 64
int launch_attn_pp(const void* q, const void* k, const void* v, void* o, int batch, int heads, int n_q, int n_kv,
                   int d, long long ld_q, long long ld_k, long long ld_v, long long ld_o, float scale,
                   const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2, int order, cudaStream_t st);

}  // namespace vf
