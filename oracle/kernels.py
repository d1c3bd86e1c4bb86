"""numpy restatement of the four per-step kernels of the VFace hot path.  TEST INFRASTRUCTURE ONLY.

Each function cites the reference lines it follows (paths relative to
/root/reference/REFace).  Arithmetic that the reference delegates to PyTorch
(torch.fft, F.grid_sample, softmax) is restated from the published semantics of
those ops (reference pin torch==1.13.1, REFace/setup.sh:3; behaviour unchanged in
torch 2.11, which is what tests/golden/ was generated with).

Pinned by tests/test_oracle_vs_golden.py (committed reference outputs) and, in
the build container, tests/test_oracle_vs_reference.py (live reference import).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------------------
# FSAI -- Frequency Spectrum Attention Interpolation
# ----------------------------------------------------------------------------------------------
def fsai_blend(donor: np.ndarray, dst: np.ndarray, split_ratio: float = 0.8) -> np.ndarray:
    """combine_fft_high_low(q1=donor, q2=dst, split_ratio)  (scripts/face_swap_utils.py:425-464).

    1-D FFT along the LAST (channel) axis; bins [0, sp) come from `dst`, bins
    [sp, d) from `donor`, sp = int(d * split_ratio); inverse FFT; keep the real
    part (the merged spectrum is not Hermitian, so .real is load-bearing).
    """
    q1 = np.asarray(donor, dtype=F32)
    q2 = np.asarray(dst, dtype=F32)
    d = q1.shape[-1]
    sp = int(d * split_ratio)
    f1 = np.fft.fft(q1.astype(np.float64), axis=-1)
    f2 = np.fft.fft(q2.astype(np.float64), axis=-1)
    comb = np.zeros_like(f1)
    comb[..., :sp] = f2[..., :sp]
    comb[..., sp:] = f1[..., sp:]
    return np.fft.ifft(comb, axis=-1).real.astype(F32)


def fsai_filter_response(d: int, split: int) -> np.ndarray:
    """Second, independent form: the op is linear and real,
    out = dst + filt(donor - dst) with the real symmetric frequency response
    h[k] = 0.5 * ([k >= split] + [(d - k) % d >= split]).  The CUDA kernel uses this form."""
    k = np.arange(d)
    return 0.5 * ((k >= split).astype(np.float64) + (((d - k) % d) >= split).astype(np.float64))


def fsai_blend_filter_form(donor, dst, split_ratio=0.8):
    q1 = np.asarray(donor, dtype=np.float64)
    q2 = np.asarray(dst, dtype=np.float64)
    d = q1.shape[-1]
    h = fsai_filter_response(d, int(d * split_ratio))
    y = np.fft.ifft(np.fft.fft(q1 - q2, axis=-1) * h, axis=-1)
    return (q2 + y.real).astype(F32)


# ----------------------------------------------------------------------------------------------
# Flow-guided warp + blend
# ----------------------------------------------------------------------------------------------
def flow_taps(flow: np.ndarray, h: int, w: int, unnormalize: str = "cpu"):
    """Index chain of warp_image (scripts/temporal_flow.py:40-53) followed by
    F.grid_sample(bilinear, padding_mode='border', align_corners=True).

    flow: (2, h, w) fp32, channel 0 = x displacement, 1 = y, feature-pixel units.
    Returns x0, y0 (int32 floor indices), ix, iy (clamped fp32 source coords).
    Every step is fp32 and in the reference's op order:
        v  = 2.0f * (p + f) / (size - 1) - 1.0f            temporal_flow.py:45-49
        i  = (v + 1) * ((size - 1) / 2)    ATen CPU  (GridSamplerKernel.cpp ComputeLocation)
        i  = ((v + 1) / 2) * (size - 1)    ATen CUDA (GridSampler.h grid_sampler_unnormalize)
        i  = min(size - 1, max(i, 0))      border padding
        i0 = floor(i)
    """
    flow = np.asarray(flow, dtype=F32)
    xs = np.arange(w, dtype=F32)[None, :].repeat(h, 0)
    ys = np.arange(h, dtype=F32)[:, None].repeat(w, 1)
    gx = xs + flow[0]
    gy = ys + flow[1]
    vx = F32(2.0) * gx / F32(max(w - 1, 1)) - F32(1.0)
    vy = F32(2.0) * gy / F32(max(h - 1, 1)) - F32(1.0)
    if unnormalize == "cpu":
        ix = (vx + F32(1.0)) * F32((w - 1) / 2.0)
        iy = (vy + F32(1.0)) * F32((h - 1) / 2.0)
    else:
        ix = ((vx + F32(1.0)) / F32(2.0)) * F32(w - 1)
        iy = ((vy + F32(1.0)) / F32(2.0)) * F32(h - 1)
    ix = np.minimum(F32(w - 1), np.maximum(ix, F32(0.0))).astype(F32)
    iy = np.minimum(F32(h - 1), np.maximum(iy, F32(0.0))).astype(F32)
    x0 = np.floor(ix).astype(np.int32)
    y0 = np.floor(iy).astype(np.int32)
    return x0, y0, ix, iy


def warp_tokens(prev: np.ndarray, flow: np.ndarray, h: int, w: int, unnormalize: str = "cpu") -> np.ndarray:
    """Bilinear border-padded gather of one frame in token layout (h*w, c)."""
    prev = np.asarray(prev, dtype=F32).reshape(h, w, -1)
    x0, y0, ix, iy = flow_taps(flow, h, w, unnormalize)
    x1 = x0 + 1
    y1 = y0 + 1
    fx0, fy0 = x0.astype(F32), y0.astype(F32)
    fx1, fy1 = fx0 + F32(1.0), fy0 + F32(1.0)
    w_nw = ((fx1 - ix) * (fy1 - iy))[..., None]
    w_ne = ((ix - fx0) * (fy1 - iy))[..., None]
    w_sw = ((fx1 - ix) * (iy - fy0))[..., None]
    w_se = ((ix - fx0) * (iy - fy0))[..., None]
    inx = x1 <= w - 1
    iny = y1 <= h - 1
    x1c = np.minimum(x1, w - 1)
    y1c = np.minimum(y1, h - 1)
    out = prev[y0, x0] * w_nw
    out = out + np.where(inx[..., None], prev[y0, x1c] * w_ne, F32(0))
    out = out + np.where(iny[..., None], prev[y1c, x0] * w_sw, F32(0))
    out = out + np.where((inx & iny)[..., None], prev[y1c, x1c] * w_se, F32(0))
    return out.reshape(h * w, -1).astype(F32)


def flow_warp_blend(x: np.ndarray, flow, alpha: float, h: int, w: int, prev_halo=None) -> np.ndarray:
    """align_by_flow (scripts/temporal_flow.py:222-237) on the native token layout
    (frames, h*w, c) used at the call site (ldm/models/pnp_utils.py:201-218):
        out[0] = x[0];  out[i+1] = alpha * x[i+1] + (1 - alpha) * warp(x[i], flow[i])
    The warp reads the UN-aligned input x[i] (no recurrence).

    prev_halo (h*w, c): optional predecessor of x[0] owned by the previous frame shard; then
    `flow` has one leading entry for it: out[0] = alpha*x[0] + (1-alpha)*warp(prev_halo, flow[0]).
    """
    x = np.asarray(x, dtype=F32)
    out = x.copy()
    a = F32(alpha)
    b = F32(1 - alpha)
    off = 0
    if prev_halo is not None:
        out[0] = a * x[0] + b * warp_tokens(prev_halo, flow[0], h, w)
        off = 1
    for i in range(x.shape[0] - 1):
        out[i + 1] = a * x[i + 1] + b * warp_tokens(x[i], flow[i + off], h, w)
    return out


# ----------------------------------------------------------------------------------------------
# Attention core
# ----------------------------------------------------------------------------------------------
def attention(q: np.ndarray, k: np.ndarray, v: np.ndarray, heads: int, scale: float,
              k2=None, v2=None) -> np.ndarray:
    """softmax(q k^T * scale) v per head (ldm/models/pnp_utils.py:270-286, attention.py:206-220),
    on the (batch, n, heads*d) layout the projections produce.  fp64 accumulation.
    Optional second key/value segment (k2, v2) is concatenated along the key axis (the
    BASELINE.json 'concatenated target K/V' microbench; no reference function -> unpinned)."""
    q = np.asarray(q, dtype=np.float64)
    k = np.asarray(k, dtype=np.float64)
    v = np.asarray(v, dtype=np.float64)
    if k2 is not None:
        k = np.concatenate([k, np.asarray(k2, dtype=np.float64)], axis=1)
        v = np.concatenate([v, np.asarray(v2, dtype=np.float64)], axis=1)
    b, n, c = q.shape
    d = c // heads
    out = np.empty((b, n, c), dtype=np.float64)
    for bi in range(b):
        for hi in range(heads):
            sl = slice(hi * d, (hi + 1) * d)
            s = (q[bi, :, sl] @ k[bi, :, sl].T) * scale
            s -= s.max(axis=-1, keepdims=True)
            p = np.exp(s)
            p /= p.sum(axis=-1, keepdims=True)
            out[bi, :, sl] = p @ v[bi, :, sl]
    return out.astype(F32)

# ----------------------------------------------------------------------------------------------
# projections of a transformer block (what csrc/vf_gemm3.cu fuses)
# ----------------------------------------------------------------------------------------------
def layer_norm(x: np.ndarray, gamma: np.ndarray, beta: np.ndarray, eps: float = 1e-5) -> np.ndarray:
    """nn.LayerNorm over the last axis (ldm/modules/attention.py:233-235: norm1 / norm2 / norm3;
    biased variance, eps inside the square root).  float64 arithmetic."""
    xd = x.astype(np.float64)
    mu = xd.mean(-1, keepdims=True)
    var = ((xd - mu) ** 2).mean(-1, keepdims=True)
    return (xd - mu) / np.sqrt(var + eps) * gamma.astype(np.float64) + beta.astype(np.float64)


def block_projection(x, w, bias=None, residual=None, ln=None, row_bias=None) -> np.ndarray:
    """One projection of a BasicTransformerBlock / SpatialTransformer with what surrounds it in the reference:

        y = Linear(LayerNorm(x)) + residual + row_bias[sample]

    ln = (gamma, beta, eps) or None.  Restates `self.attn1(self.norm1(x))`'s norm + to_q / to_k / to_v
    (attention.py:239, :172-174, with w = [Wq; Wk; Wv]), `to_out(...) + x` and the attn2 row of a single-token context
    (attention.py:176, :239-241), proj_in and `proj_out(x) + x_in` as 1x1 convolutions over tokens (attention.py:261-288).
    x (..., k), w (n, k), bias (n), residual (..., n), row_bias (batch, n) for x (batch, tokens, k).  float64."""
    xd = layer_norm(x, *ln) if ln is not None else x.astype(np.float64)
    y = xd @ w.astype(np.float64).T
    if bias is not None:
        y = y + bias.astype(np.float64)
    if row_bias is not None:
        y = y + row_bias.astype(np.float64)[:, None, :]
    if residual is not None:
        y = y + residual.astype(np.float64)
    return y


# ----------------------------------------------------------------------------------------------
# CFG + DDIM update
# ----------------------------------------------------------------------------------------------
def ddim_cfg_step(x, e_uncond, e_cond, a_t, a_prev, sigma_t, sqrt_one_minus_at, scale, noise=None):
    """Classifier-free guidance + DDIM update (ldm/models/diffusion/ddim_w_inv.py:666, :686, :696-700):
        e       = e_u + s (e_c - e_u)
        pred_x0 = (x - sqrt(1-a_t) e) / sqrt(a_t)
        x_prev  = sqrt(a_prev) pred_x0 + sqrt(1 - a_prev - sigma^2) e + sigma * noise
    All fp32, scalars are the fp32 table entries of the sampler."""
    x = np.asarray(x, dtype=F32)
    eu = np.asarray(e_uncond, dtype=F32)
    ec = np.asarray(e_cond, dtype=F32)
    e = eu + F32(scale) * (ec - eu)
    a_t, a_prev, sigma_t, s1m = F32(a_t), F32(a_prev), F32(sigma_t), F32(sqrt_one_minus_at)
    pred_x0 = (x - s1m * e) / np.sqrt(a_t)
    dir_xt = np.sqrt(F32(1.0) - a_prev - sigma_t * sigma_t) * e
    x_prev = np.sqrt(a_prev) * pred_x0 + dir_xt
    if noise is not None:
        x_prev = x_prev + sigma_t * np.asarray(noise, dtype=F32)
    return x_prev.astype(F32), pred_x0.astype(F32)


# ----------------------------------------------------------------------------------------------
# DDIM tables
# ----------------------------------------------------------------------------------------------
def make_schedule(S: int, eta: float = 0.0, n_ddpm: int = 1000, linear_start=0.00085, linear_end=0.012):
    """DDIM tables (ddim_w_inv.py:155-184; util.py:21-25, :46-74; betas from project_ffhq.yaml:5-9).
    The model's alphas_cumprod buffer is fp32 (ddpm.py:277 to_torch), and make_schedule indexes the
    fp32 buffer (ddim_w_inv.py:158,174), so the tables are fp32 values of the fp64 cumprod."""
    betas = np.linspace(linear_start ** 0.5, linear_end ** 0.5, n_ddpm, dtype=np.float64) ** 2
    acp = np.cumprod(1.0 - betas, axis=0).astype(F32)
    c = n_ddpm // S
    steps = np.asarray(list(range(0, n_ddpm, c))) + 1
    alphas = acp[steps]
    alphas_prev = np.asarray([acp[0]] + acp[steps[:-1]].tolist(), dtype=F32)
    # util.py:70 mixes a fp32 torch tensor (alphas) with a fp64 numpy array (alphas_prev).  numpy defers
    # `ndarray / Tensor` to Tensor.__rtruediv__, which torch evaluates as reciprocal(self) * other, so
    # 1/(1 - alphas) is formed in fp32 and only then promoted; every other term is fp64.
    a64, ap64 = alphas.astype(np.float64), alphas_prev.astype(np.float64)
    recip = (F32(1.0) / (F32(1.0) - alphas)).astype(np.float64)
    sigmas = eta * np.sqrt((recip * (1 - ap64)) * (1 - a64 / ap64))
    return dict(ddim_timesteps=steps, ddim_alphas=alphas, ddim_alphas_prev=alphas_prev,
                ddim_sigmas=sigmas.astype(F32),
                ddim_sqrt_one_minus_alphas=np.sqrt(F32(1.0) - alphas).astype(F32),
                alphas_cumprod=acp)
