"""CPU restatement (PyTorch fp32) of the whole VFace denoising hot path.  TEST INFRASTRUCTURE ONLY.

The reference path is plain PyTorch, so the port is plain PyTorch too, but written functionally over
a state dict (no nn.Module tree) and with the reference's materialised N x N attention, per-pair
grid_sample loop and torch.fft calls -- i.e. it computes what the reference computes, the way the
reference computes it, on the host cores.  It is the checker for the GPU path on the GPU box (where
/root/reference does not exist) and the `cpu_baseline` / `--impl reference` arm of bench.py
(kind "port").

Reference lines followed (relative to /root/reference/REFace):
  UNet forward            ldm/modules/diffusionmodules/openaimodel.py:860-907, ResBlock :255-275,
                          Downsample/Upsample :91-160, GroupNorm32 util.py:214-216, timestep_embedding :151-171
  transformer blocks      ldm/modules/attention.py:236-243, :278-288, GEGLU :37-45
  patched attn1 + hooks   ldm/models/pnp_utils.py:92-288 (fusion "replace", "fft", "flow_fix")
  FSAI                    scripts/face_swap_utils.py:425-464
  flow warp               scripts/temporal_flow.py:40-53, :222-237
  sampler step            ldm/models/diffusion/ddim_w_inv.py:621-738 (and :564-617, :360-490)
  schedule                ldm/models/diffusion/ddim_w_inv.py:155-184, util.py:21-25, :46-74

Pinned against the imported reference by tests/test_oracle_vs_reference.py (build container) and
against committed reference outputs by tests/test_oracle_vs_golden.py.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import kernels as ok

SD = Dict[str, torch.Tensor]


# ---- primitive layers ---------------------------------------------------------------------------
def _gn(sd: SD, p: str, x, eps):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps).type(x.dtype)


def _conv(sd: SD, p: str, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def _lin(sd: SD, p: str, x):
    return F.linear(x, sd[p + ".weight"], sd.get(p + ".bias"))


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def resblock(sd: SD, p: str, x, emb):
    h = _conv(sd, p + ".in_layers.2", F.silu(_gn(sd, p + ".in_layers.0", x, 1e-5)))
    h = h + _lin(sd, p + ".emb_layers.1", F.silu(emb))[:, :, None, None]
    h = _conv(sd, p + ".out_layers.3", F.silu(_gn(sd, p + ".out_layers.0", h, 1e-5)))
    if p + ".skip_connection.weight" in sd:
        w = sd[p + ".skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + ".skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


# ---- VFace hooks (reference formulation) ------------------------------------------------------------
def combine_fft_high_low(q1, q2, split_ratio):
    q1, q2 = q1.float(), q2.float()
    f1, f2 = torch.fft.fft(q1, dim=-1), torch.fft.fft(q2, dim=-1)
    sp = int(q1.size(-1) * split_ratio)
    comb = torch.zeros_like(f1)
    comb[..., :sp] = f2[..., :sp]
    comb[..., sp:] = f1[..., sp:]
    return torch.fft.ifft(comb, dim=-1).real.to(torch.float32)


def warp_image(img, flow):
    _, _, h, w = img.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    vgrid = torch.stack([xs, ys])[None] + flow
    vgrid[:, 0] = 2.0 * vgrid[:, 0] / max(w - 1, 1) - 1.0
    vgrid[:, 1] = 2.0 * vgrid[:, 1] / max(h - 1, 1) - 1.0
    return F.grid_sample(img, vgrid.permute(0, 2, 3, 1), align_corners=True, padding_mode="border")


def align_by_flow(x, flow, alpha):
    out = x.clone()
    for i in range(x.shape[0] - 1):
        warped = warp_image(x[i].unsqueeze(0), flow[i])
        out[i + 1] = alpha * x[i + 1] + (1 - alpha) * warped.squeeze(0)
    return out


def apply_hooks(q, k, hook: Optional[dict]):
    """In-place rewrite of q/k chunks [uncond ; cond ; recon]; chunk 0 is the donor (pnp_utils.py:131-222)."""
    if not hook or not hook.get("switch_on", False):
        return
    B = q.shape[0] // 3
    fusion, sr = hook["fusion"], hook.get("split_ratio_fft", 0.8)
    if fusion == "replace":
        for t in (q, k):
            t[B:2 * B] = t[:B]
            t[2 * B:] = t[:B]
        return
    if fusion not in ("fft", "flow_fix"):
        raise NotImplementedError(fusion)
    q[B:2 * B] = combine_fft_high_low(q[:B], q[B:2 * B], sr)
    k[B:2 * B] = combine_fft_high_low(k[:B], k[B:2 * B], sr)
    q[2 * B:] = combine_fft_high_low(q[:B], q[2 * B:], sr)
    k[2 * B:] = combine_fft_high_low(k[:B], k[2 * B:], sr)
    flow = hook.get("flow")
    if fusion == "flow_fix" and flow is not None and q.shape[1] == 4096:
        for t in (q, k):
            f = t[B:2 * B].reshape(B, 64, 64, -1).permute(0, 3, 1, 2)
            f = align_by_flow(f, flow, hook.get("alpha", 0.8))
            t[B:2 * B] = f.permute(0, 2, 3, 1).reshape(-1, 64 * 64, f.shape[1])


def attention_core(q, k, v, heads):
    """Materialised softmax(q k^T scale) v, head by head to bound host memory (pnp_utils.py:270-286)."""
    b, n, c = q.shape
    d = c // heads
    scale = d ** -0.5
    out = torch.empty_like(q)
    for h in range(heads):
        sl = slice(h * d, (h + 1) * d)
        sim = torch.einsum("bid,bjd->bij", q[..., sl], k[..., sl]) * scale
        out[..., sl] = torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v[..., sl])
    return out


def cross_attention(sd: SD, p: str, x, context, heads, hook=None):
    q = _lin(sd, p + ".to_q", x)
    ctx = x if context is None else context
    k = _lin(sd, p + ".to_k", ctx)
    v = _lin(sd, p + ".to_v", ctx)
    if context is None:
        apply_hooks(q, k, hook)
    return _lin(sd, p + ".to_out.0", attention_core(q, k, v, heads))


def spatial_transformer(sd: SD, p: str, x, context, heads, hook=None):
    b, c, h, w = x.shape
    x_in = x
    x = _conv(sd, p + ".proj_in", F.group_norm(x, 32, sd[p + ".norm.weight"], sd[p + ".norm.bias"], 1e-6), padding=0)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)
    t = p + ".transformer_blocks.0"
    ln = lambda name, z: F.layer_norm(z, (z.shape[-1],), sd[f"{t}.{name}.weight"], sd[f"{t}.{name}.bias"])
    x = cross_attention(sd, t + ".attn1", ln("norm1", x), None, heads, hook) + x
    x = cross_attention(sd, t + ".attn2", ln("norm2", x), context, heads) + x
    g = _lin(sd, t + ".ff.net.0.proj", ln("norm3", x))
    a, gate = g.chunk(2, dim=-1)
    x = _lin(sd, t + ".ff.net.2", a * F.gelu(gate)) + x
    x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)
    return _conv(sd, p + ".proj_out", x, padding=0) + x_in


# ---- UNet -------------------------------------------------------------------------------------------
def unet_forward(sd: SD, x, timesteps, context, heads: int, hooks: Optional[Dict[str, dict]] = None):
    """Structure is read off the state-dict keys.  `hooks` maps an attn1 module name prefix
    (e.g. 'input_blocks.1.1') to its hook configuration."""
    hooks = hooks or {}
    mc = sd["time_embed.0.weight"].shape[1]
    emb = _lin(sd, "time_embed.2", F.silu(_lin(sd, "time_embed.0", timestep_embedding(timesteps, mc))))

    def block(prefix, h):
        j = 0
        while any(key.startswith(f"{prefix}.{j}.") for key in keyset):
            p = f"{prefix}.{j}"
            if p + ".in_layers.0.weight" in sd:
                h = resblock(sd, p, h, emb)
            elif p + ".norm.weight" in sd:
                h = spatial_transformer(sd, p, h, context, heads, hooks.get(p))
            elif p + ".op.weight" in sd:
                h = _conv(sd, p + ".op", h, stride=2)
            elif p + ".conv.weight" in sd:
                h = _conv(sd, p + ".conv", F.interpolate(h, scale_factor=2, mode="nearest"))
            else:
                h = _conv(sd, p, h)
            j += 1
        return h

    keyset = list(sd.keys())
    n_in = 1 + max(int(k.split(".")[1]) for k in keyset if k.startswith("input_blocks."))
    n_out = 1 + max(int(k.split(".")[1]) for k in keyset if k.startswith("output_blocks."))
    h = x.float()
    hs = []
    for i in range(n_in):
        h = block(f"input_blocks.{i}", h)
        hs.append(h)
    h = block("middle_block", h)
    for i in range(n_out):
        h = block(f"output_blocks.{i}", torch.cat([h, hs.pop()], dim=1))
    h = _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h, 1e-5)))
    return h


def vface_hooks(sd: SD, flow, fusion="flow_fix", split_ratio_fft=0.8, alpha=0.8, where=("input_blocks",)):
    """Effective configuration of ddim_w_inv.py:300-308: switched on for the input-block attn1 modules."""
    names = sorted({k.split(".transformer_blocks")[0] for k in sd
                    if k.endswith(".transformer_blocks.0.attn1.to_q.weight") and k.split(".")[0] in where})
    return {n: dict(switch_on=True, fusion=fusion, split_ratio_fft=split_ratio_fft, alpha=alpha, flow=flow) for n in names}


# ---- sampler ------------------------------------------------------------------------------------------
@torch.no_grad()
def sample(sd: SD, heads: int, S: int, x_T, cond, target_cond, uc, inpaint_image, inpaint_mask, inversion: dict,
           flow, scale: float = 3.0, eta: float = 0.0, hooks_fusion="flow_fix", return_all=False, max_steps=None,
           step_callback=None):
    """DDIMSampler.sample -> ddim_sampling -> p_sample_ddim_with_inverse (ddim_w_inv.py:186-355, :621-738); noise comes
    from the global CPU generator exactly as in the reference (two draws per step)."""
    tb = ok.make_schedule(S, eta)
    steps = tb["ddim_timesteps"]
    hooks = vface_hooks(sd, flow, fusion=hooks_fusion) if hooks_fusion else {}
    img = x_T.float()
    b = img.shape[0]
    extra = torch.cat([inpaint_image, inpaint_mask], dim=1).float()
    xs, x0s = [], []
    for i, step in enumerate(np.flip(steps)):
        if max_steps is not None and i >= max_steps:
            break
        index = len(steps) - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        x_full = torch.cat([img, extra], dim=1)
        inv_full = torch.cat([inversion[int(step)].float(), extra], dim=1)
        x_in = torch.cat([x_full, x_full, inv_full], dim=0)
        c_in = torch.cat([uc, cond, target_cond], dim=0).float()
        e_u, e_c, _ = unet_forward(sd, x_in, torch.cat([ts] * 3), c_in, heads, hooks).chunk(3)
        # noise_like is called TWICE per step, whatever eta is (ddim_w_inv.py:697 for x_prev, :704 for the discarded recon
        # branch): the first draw is the noise of x_prev, the second only advances the generator
        noise = torch.randn(img.shape).numpy()
        torch.randn(img.shape)
        if eta == 0:
            noise = None
        x_prev, pred_x0 = ok.ddim_cfg_step(img.numpy(), e_u.numpy(), e_c.numpy(), tb["ddim_alphas"][index],
                                           tb["ddim_alphas_prev"][index], tb["ddim_sigmas"][index],
                                           tb["ddim_sqrt_one_minus_alphas"][index], scale, noise)
        img = torch.from_numpy(x_prev)
        xs.append(img)
        x0s.append(torch.from_numpy(pred_x0))
        if step_callback:
            step_callback(i, img)
    return (img, xs, x0s) if return_all else img


@torch.no_grad()
def p_sample_ddim(sd: SD, heads: int, S: int, index: int, step: int, x, cond, uc, inpaint_image, inpaint_mask,
                  scale: float = 3.0, eta: float = 0.0):
    """DDIMSampler.p_sample_ddim (ddim_w_inv.py:564-617): the 2-way [uncond ; cond] step on an un-hooked UNet (one branch
    when scale == 1, :577-578); ONE noise draw (:611)."""
    tb = ok.make_schedule(S, eta)
    b = x.shape[0]
    ts = torch.full((b,), int(step), dtype=torch.long)
    x_full = torch.cat([x.float(), inpaint_image.float(), inpaint_mask.float()], dim=1)
    if uc is None or scale == 1.0:
        e_c = unet_forward(sd, x_full, ts, cond.float(), heads, {})
        e_u, scale = e_c, 1.0
    else:
        e_u, e_c = unet_forward(sd, torch.cat([x_full] * 2), torch.cat([ts] * 2), torch.cat([uc, cond]).float(), heads, {}).chunk(2)
    noise = torch.randn(x.shape).numpy()
    x_prev, pred_x0 = ok.ddim_cfg_step(x.float().numpy(), e_u.numpy(), e_c.numpy(), tb["ddim_alphas"][index], tb["ddim_alphas_prev"][index],
                                       tb["ddim_sigmas"][index], tb["ddim_sqrt_one_minus_alphas"][index], scale,
                                       noise if eta > 0 else None)
    return torch.from_numpy(x_prev), torch.from_numpy(pred_x0)
