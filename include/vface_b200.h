/*
 * vface_b200 -- C-ABI of the B200-native VFace denoising hot path.
 *
 * The reference (Sanoojan/VFace, REFace/) is pure Python/PyTorch and has no FFI
 * of its own, so every entry point below replaces a
 * span of reference Python; the file:line each one replaces is cited.  Paths are
 * relative to /root/reference/REFace.
 *
 * Conventions:
 *   - the caller owns every buffer; nothing here allocates device memory;
 *   - every launch is asynchronous on `stream` (a cudaStream_t passed as void*)
 *     and is CUDA-graph capturable;
 *   - return 0 on success, non-zero on failure with a message in
 *     vf_last_error() (thread-local);
 *   - there is no CPU fallback: without an sm_100a device the calls fail.
 *   - one process drives ONE GPU: every launch goes to the CURRENT CUDA device, and
 *     the library binds itself to the device that is current at the first call (its
 *     function attributes, cuBLASLt handle/plans and grid sizes are per device).  A
 *     later call with another device current fails with a message instead of
 *     launching there.  Multi-GPU = one process per GPU (torch.distributed / torchrun).
 *   - dtype: VF_F32 is the fp32 reference-precision path, VF_BF16 the tensor-core
 *     path (bf16 storage, fp32 accumulation).
 *   - token tensors use the reference's native (batch, n, heads*d_head) layout;
 *     `ld_*` are row strides in ELEMENTS (>= heads*d_head), batch stride is n*ld.
 */
#ifndef VFACE_B200_H_
#define VFACE_B200_H_

#ifdef __cplusplus
extern "C" {
#endif

#define VF_F32 0
#define VF_BF16 1

#define VF_ABI_VERSION 10

/* ABI version of the loaded library (== VF_ABI_VERSION). */
int vf_abi_version(void);

/* Message of the last failing call on this thread ("" if none). */
const char* vf_last_error(void);

/*
 * Fused attention core: o = softmax(q k^T * scale) v per (batch, head), streaming
 * (no n_q x n_kv matrix in HBM).
 * Replaces ldm/models/pnp_utils.py:270-286 (patched attn1) and
 * ldm/modules/attention.py:203-220 (CrossAttention.forward core).
 * Optional second key/value segment (k2, v2, n_kv2 rows) is treated as concatenated
 * after (k, v) along the key axis ("injected target-frame K/V"; BASELINE.json config 3;
 * no reference function).  Pass k2 = v2 = NULL, n_kv2 = 0 to disable.
 * d_head must be a multiple of 8: 8 <= d_head <= 192, or one of 256 / 384 / 512 (wide heads: the first-stage AttnBlock,
 * ldm/modules/diffusionmodules/model.py:150-203, is one head of width 512) in bf16; <= 256 in fp32.
 */
int vf_attn_fwd(const void* q, const void* k, const void* v, void* o,
                int batch, int heads, int n_q, int n_kv, int d_head,
                long long ld_q, long long ld_k, long long ld_v, long long ld_o,
                float scale,
                const void* k2, const void* v2, int n_kv2, long long ld_k2, long long ld_v2,
                int dtype, void* stream);

/*
 * Frequency Spectrum Attention Interpolation of one (donor, dst) pair:
 *   out[r, :] = Re ifft( [ fft(dst[r])[0:split] , fft(donor[r])[split:d] ] )
 * along the channel axis (length d) of each of `rows` token rows.
 * Replaces scripts/face_swap_utils.py:425-464 (combine_fft_high_low) and the slice
 * assignment at ldm/models/pnp_utils.py:179-183 / :195-199.  `out` may alias `dst`
 * (the reference assigns in place).  d must be 2^a * 5^b with 32 <= d <= 2048.
 */
int vf_fsai_blend(const void* donor, const void* dst, void* out,
                  long long rows, int d, int split,
                  long long ld_donor, long long ld_dst, long long ld_out,
                  int dtype, void* stream);

/*
 * The two calls of one tensor (q or k) of a hooked module fused: both branches share the
 * donor (chunk 0).  out_a/out_b may alias dst_a/dst_b.
 * Replaces ldm/models/pnp_utils.py:195+:198 (q) or :196+:199 (k).
 */
int vf_fsai_blend2(const void* donor,
                   const void* dst_a, void* out_a,
                   const void* dst_b, void* out_b,
                   long long rows, int d, int split,
                   long long ld_donor, long long ld_a, long long ld_out_a,
                   long long ld_b, long long ld_out_b,
                   int dtype, void* stream);

/*
 * Flow-Guided Attention Temporal Smoothening on the native token layout
 * x: (frames, h*w, c), row stride ld_x:
 *   out[0]   = x[0]                                     (or blended with prev_halo, below)
 *   out[i+1] = alpha * x[i+1] + (1-alpha) * bilinear_border_sample(x[i], p + flow[i])
 * Replaces scripts/temporal_flow.py:222-237 (align_by_flow) + :40-53 (warp_image) and the
 * permute/reshape round trip at ldm/models/pnp_utils.py:206-218.
 * flow: fp32 (n_flow, 2, h, w), channel 0 = x displacement, 1 = y, feature-pixel units.
 *   prev_halo == NULL: n_flow = frames-1, flow[i] maps frame i -> i+1.
 *   prev_halo != NULL: (h*w, c) rows (stride ld_halo) = last frame of the previous frame
 *     shard; n_flow = frames and flow[0] maps the halo frame -> frame 0.
 * `out` must not alias `x`.  The source index chain is fp32 in the reference's op order
 * (floor indices are bit-exact with torch grid_sample).  c*sizeof(elem) must be a multiple of 16.
 * taps_out (optional, may be NULL): int32 (n_flow, h*w, 2) receives the (x0, y0) floor
 * indices used, for index-parity tests.
 * alpha is a double so that (1 - alpha) is formed in double like the reference's Python scalar
 * (temporal_flow.py:234) before both weights are rounded to fp32.
 */
int vf_flow_warp_blend(const void* x, const void* prev_halo, const float* flow, void* out,
                       int frames, int h, int w, int c,
                       long long ld_x, long long ld_halo, long long ld_out,
                       double alpha, int dtype, int* taps_out, void* stream);

/*
 * Classifier-free guidance + DDIM update, fused:
 *   e = e_u + s (e_c - e_u); pred_x0 = (x - sqrt(1-a_t) e)/sqrt(a_t);
 *   x_prev = sqrt(a_prev) pred_x0 + sqrt(1 - a_prev - sigma^2) e + sigma * noise
 * Replaces ldm/models/diffusion/ddim_w_inv.py:666 and :679-700 (also :591, :604-616).
 * x, x_prev, pred_x0, noise are fp32 (n elements each); e_uncond / e_cond have
 * dtype `dtype_e`.  noise may be NULL when sigma_t == 0.
 */
int vf_ddim_cfg_step(const void* x, const void* e_uncond, const void* e_cond,
                     void* x_prev, void* pred_x0,
                     float a_t, float a_prev, float sigma_t, float sqrt_one_minus_at,
                     float cfg_scale, const void* noise, long long n, int dtype_e, void* stream);

/*
 * Forward-DDIM (inversion) update of ddim_invert, fused with optional CFG:
 *   e = e_c (e_uncond == NULL) or e_u + s (e_c - e_u)
 *   x_next = (x - sqrt(1-a_cur) e) * sqrt(a_next)/sqrt(a_cur) + sqrt(1-a_next) e
 * Replaces ldm/models/diffusion/ddim_w_inv.py:426-449.
 */
int vf_ddim_invert_step(const void* x, const void* e_uncond, const void* e_cond, void* x_next,
                        float a_cur, float a_next, float cfg_scale,
                        long long n, int dtype_e, void* stream);

/*
 * ---- fused glue kernels of the UNet between its GEMM-shaped ops ----------
 * All take channels-last / token-major activations: rows of `c` contiguous channels.
 */

/*
 * GroupNorm32 on channels-last activations x: (n, hw, c), fp32 statistics:
 *   y = act( GN(x + add_nc[n, :]) * gamma + beta ),  act = SiLU if silu != 0.
 * Replaces normalization() -> SiLU at ldm/modules/diffusionmodules/openaimodel.py:201-205, :236-239
 * (GroupNorm32, util.py:214-216), the `h + emb_out` add of ResBlock._forward (:265-273, through
 * add_nc, which may be NULL) and Normalize() of SpatialTransformer (attention.py:278-281).
 * workspace: vf_group_norm_workspace_floats(n, hw, groups) floats of device scratch.
 */
long long vf_group_norm_workspace_floats(int n, int hw, int groups);
int vf_group_norm_nhwc(const void* x, const void* add_nc, const void* gamma, const void* beta, void* y,
                       float* workspace, int n, int hw, int c, int groups, float eps, int silu,
                       int dtype, void* stream);

/*
 * The same over the channel concatenation [x (c1 channels) ; x2 (c2 channels)] of two channels-last tensors,
 * read in place: y (n, hw, c1+c2) = act(GN(cat(x, x2) + add_nc)).  Replaces `th.cat([h, hs.pop()], dim=1)`
 * (openaimodel.py:899) followed by the first GroupNorm of the output-block ResBlock (:201-205): the
 * concatenated skip tensor is never materialised.  x2 == NULL reduces to vf_group_norm_nhwc.
 */
int vf_group_norm_nhwc_cat(const void* x, int c1, const void* x2, int c2, const void* add_nc,
                           const void* gamma, const void* beta, void* y, float* workspace,
                           int n, int hw, int groups, float eps, int silu, int dtype, void* stream);

/*
 * res = x + y + bias ; out = LayerNorm(res) * gamma + beta over the last axis (c).
 * y and bias may be NULL (then res may be NULL and out = LayerNorm(x)).  bias is (rows/rows_per_bias, c)
 * when rows_per_bias > 0 (one row per sample) or a single (c) row when rows_per_bias == 0.
 * Replaces `attn(norm(x)) + x` followed by the next nn.LayerNorm in
 * BasicTransformerBlock._forward (ldm/modules/attention.py:239-243).
 */
int vf_add_layer_norm(const void* x, const void* y, const void* bias, long long rows_per_bias,
                      const void* gamma, const void* beta, void* res, void* out,
                      long long rows, int c, float eps, int dtype, void* stream);

/*
 * GEGLU gate: out[r, 0:k] = h[r, 0:k] * gelu(h[r, k:2k]) (exact erf GELU); h row stride ld_h >= 2k.
 * Replaces GEGLU.forward (ldm/modules/attention.py:43-45).
 */
int vf_geglu(const void* h, void* out, long long rows, int k, long long ld_h, int dtype, void* stream);

/*
 * Feed-forward up-projection with the GEGLU gate fused into the GEMM epilogue (tcgen05 tensor cores):
 *   out[r, 0:n] = (x[r] . Wv^T + bv) * gelu(x[r] . Wg^T + bg),  w = [Wv ; Wg] (2n, k) row-major, bias (2n) or NULL.
 * Replaces GEGLU.forward (ldm/modules/attention.py:37-45: proj -> chunk(2) -> x * gelu(gate)); the (rows, 2n)
 * projection is never written to memory.  bf16 only (the fp32 path keeps library GEMM + vf_geglu);
 * k % 8 == 0, n % 128 == 0; x row stride ld_x elements, out is (rows, n) contiguous.
 */
int vf_linear_geglu(const void* x, const void* w, const void* bias, void* out,
                    long long rows, int k, int n, long long ld_x, int dtype, void* stream);

/*
 * out = a + b + bias (b, bias optional; bias rows as in vf_add_layer_norm).  Residual adds and the
 * convolution biases of ResBlock / SpatialTransformer (openaimodel.py:275, attention.py:288).
 */
int vf_add_bias(const void* a, const void* b, const void* bias, long long rows_per_bias, void* out,
                long long rows, int c, int dtype, void* stream);

/*
 * Nearest-neighbour 2x upsampling of a channels-last map x: (n, h, w, c) -> out: (n, 2h, 2w, c).
 * Replaces F.interpolate(x, scale_factor=2, mode="nearest") in Upsample.forward
 * (ldm/modules/diffusionmodules/openaimodel.py:107-117; also model.py:52-56 of the first-stage decoder).
 */
int vf_upsample_nearest2x_nhwc(const void* x, void* out, int n, int h, int w, int c, int dtype, void* stream);

/*
 * Output convolution of the UNet with an fp32 result: eps = conv3x3(x, weight, padding 1) + bias,
 *   x (n, h, w, c) channels-last bf16, weight (c_out = 4, c, 3, 3) contiguous OIHW bf16, bias (4) bf16 or NULL,
 *   out (n, 4, h, w) contiguous fp32.
 * Replaces the last layer of `self.out` (ldm/modules/diffusionmodules/openaimodel.py:835, applied at :907) on the bf16
 * path: the classifier-free-guidance combination (ldm/models/diffusion/ddim_w_inv.py:666) multiplies the rounding of a
 * bf16 eps by 3.6, so eps leaves the UNet in fp32.  The fp32 path keeps the library convolution.
 */
int vf_conv3x3_out_f32(const void* x, const void* weight, const void* bias, void* out, int n, int h, int w, int c,
                       int c_out, int dtype, void* stream);

/*
 * Projection with bias and residual add inside one library GEMM (cuBLASLt, beta = 1, bias epilogue):
 *   out[r, 0:n] = residual[r, 0:n] + x[r, 0:k] . W^T + bias,   W (n, k) row-major, bias (n) or NULL.
 * Replaces `x = ff(norm3(x)) + x` (ldm/modules/attention.py:242: FeedForward's down-projection, :61-64, plus the add)
 * and `proj_out(x) + x_in` (attention.py:287-288).  out must not alias residual (measured: the library's in-place
 * path rounds differently by one bf16 ulp).  Row strides in elements; the caller owns
 * `workspace` (workspace_bytes may be 0).  dtype VF_BF16 or VF_F32 (fp32 accumulation either way).
 */
int vf_linear_residual(const void* x, const void* w, const void* bias, const void* residual, void* out,
                       long long rows, int k, int n, long long ld_x, long long ld_res, long long ld_out,
                       void* workspace, long long workspace_bytes, int dtype, void* stream);

/*
 * The same with a bias per batch entry: x, residual, out hold `batch` entries of `rows` rows each (contiguous per
 * entry), bias is (batch, n).  Replaces `x = attn1(norm1(x)) + x ; x = attn2(norm2(x), ctx) + x`
 * (ldm/modules/attention.py:240-241) when the context is a single token: attn2's output is then one row per sample,
 * which rides in the bias of the to_out projection of attn1 (to_out: attention.py:173-176).
 */
int vf_linear_residual_batched(const void* x, const void* w, const void* bias, const void* residual, void* out,
                               int batch, long long rows, int k, int n, long long ld_x, long long ld_res, long long ld_out,
                               void* workspace, long long workspace_bytes, int dtype, void* stream);

/*
 * Projection of the 64x64 transformer level as ONE tcgen05 kernel, with the LayerNorm in front of it, the bias (one row,
 * or one row per sample) and the residual add behind it folded in:
 *   out[r, 0:n] = LN(x[r, 0:k]) . W^T + bias[r / rows_per_bias, 0:n] + residual[r, 0:n]
 * Replaces `self.attn1(self.norm1(x))`'s norm + to_q / to_k / to_v (ldm/modules/attention.py:239, :172-174; the hooked
 * closure's projections, ldm/models/pnp_utils.py:106,127-128), to_out + the adds of attention.py:239-241 (:176,:221;
 * pnp_utils.py:287), and the 1x1 convolutions proj_in / proj_out (+ x_in) of SpatialTransformer (attention.py:261-288).
 *   w          (n, k) bf16 row-major.  LayerNorm form: the caller passes W o gamma (gamma folded into the columns, rounded
 *              to bf16) and ln_colsum[j] = sum_k float(w[j, k]) (fp32, n entries); LN's beta enters through the bias
 *              (beta . W^T); mean and rstd are computed in the kernel from the raw rows (eps = ln_eps).
 *              ln_colsum NULL: no normalisation.
 *   bias       fp32, (n) with rows_per_bias = 0, or (rows / rows_per_bias, n) with rows_per_bias % 128 == 0; may be NULL.
 *   residual   bf16 (rows, n), row stride ld_res, or NULL (not together with ln_colsum).  out (rows, n) bf16, row stride
 *              ld_out; no aliasing.
 *   stats_out  optional (rows, n / 160, 2) fp32: per row and 160-column slice, {sum, sum of squares} of the bf16 OUTPUT
 *              values (not in the LayerNorm form).  ln_stats_in: the same array written by the call that produced x
 *              ((rows, ln_stats_parts, 2), partials summed here): mean / rstd then come from it instead of a second look
 *              at x -- the LayerNorm statistics are a by-product of the producer's epilogue.  NULL: computed in-kernel.
 * k a multiple of 64 up to 320, n a multiple of 160 (vf_linear_proj_supported), bf16 only; other shapes and fp32 stay on
 * the library GEMM (vf_linear_residual).
 */
int vf_linear_proj_supported(long long rows, int k, int n);
int vf_linear_proj(const void* x, const void* w, const float* bias, long long rows_per_bias, const void* residual,
                   const float* ln_colsum, float ln_eps, const float* ln_stats_in, int ln_stats_parts, float* stats_out,
                   void* out, long long rows, int k, int n, long long ld_x, long long ld_res, long long ld_out, int dtype,
                   void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VFACE_B200_H_ */
