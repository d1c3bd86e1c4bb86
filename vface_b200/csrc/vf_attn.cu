// vf_attn_fwd: argument checking and dispatch between the tcgen05 (bf16) and fp32 kernels.
#include "vf_attn.cuh"


extern "C" int vf_attn_fwd(const void* q, const void* k, const void* v, void* o,
                           int batch, int heads, int n_q, int n_kv, int d_head,
                           long long ld_q, long long ld_k, long long ld_v, long long ld_o,
                           float scale,
                           const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2,
                           int dtype, void* stream) {
  using namespace vf;
  if (int rc = check_device()) return rc;
  if (!q || !k || !v || !o) return fail("vf_attn_fwd: null pointer");
  if (batch <= 0 || heads <= 0 || n_q <= 0 || n_kv <= 0 || d_head <= 0)
    return fail("vf_attn_fwd: bad shape batch=%d heads=%d n_q=%d n_kv=%d d=%d", batch, heads, n_q, n_kv, d_head);
  if ((long long)batch * heads > 65535) return fail("vf_attn_fwd: batch*heads=%lld exceeds 65535", (long long)batch * heads);
  if (n_kv2 < 0 || ((k2 == nullptr) != (v2 == nullptr))) return fail("vf_attn_fwd: k2/v2/n_kv2 inconsistent");
  if (!(scale > 0.0f)) return fail("vf_attn_fwd: scale must be positive");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == VF_BF16)
    return launch_attn_tc(q, k, v, o, batch, heads, n_q, n_kv, d_head, ld_q, ld_k, ld_v, ld_o, scale,
                          k2, v2, n_kv2, ld_k2, ld_v2, st);
  if (dtype != VF_F32) return fail("vf_attn_fwd: bad dtype %d", dtype);
  if (d_head % 4 || d_head > 256) return fail("vf_attn_fwd(fp32): d_head=%d must be a multiple of 4, <= 256", d_head);
  AttnF32Params P;
  P.q = (const float*)q; P.k = (const float*)k; P.v = (const float*)v; P.o = (float*)o;
  P.k2 = (k2 && n_kv2 > 0) ? (const float*)k2 : nullptr;
  P.v2 = (k2 && n_kv2 > 0) ? (const float*)v2 : nullptr;
  P.heads = heads; P.n_q = n_q; P.n_kv = n_kv; P.n_kv2 = P.k2 ? n_kv2 : 0; P.d = d_head;
  P.ld_q = ld_q; P.ld_k = ld_k; P.ld_v = ld_v; P.ld_o = ld_o; P.ld_k2 = ld_k2; P.ld_v2 = ld_v2;
  P.scale = scale;
  return launch_attn_f32(P, batch, st);
}
