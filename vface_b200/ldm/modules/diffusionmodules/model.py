"""First-stage (VAE) decoder on the vface_b200 kernels -- SURVEY.md section 8(f) row 4, the step after the path.

Host-side mirror of REFace/ldm/modules/diffusionmodules/model.py: Decoder :462-570, ResnetBlock :82-147,
AttnBlock :156-203, Upsample :42-58, and of AutoencoderKL.decode (ldm/models/autoencoder.py:330-333).
Module tree and state-dict keys are the reference's (conv_in, mid.block_1, mid.attn_1, mid.block_2,
up.{level}.block.{i}, up.{level}.upsample.conv, norm_out, conv_out), so `first_stage_model.*` of the
reference checkpoint loads as is; constructor keywords are the reference's `ddconfig`
(configs/project_ffhq.yaml: ch 128, ch_mult (1,2,4,4), 2 res blocks, z_channels 4, no attn_resolutions).

Execution plan: channels-last activations in the parameter dtype, GroupNorm+SiLU as one vface_b200 kernel
(vf_group_norm_nhwc), residual add + conv bias as one (vf_add_bias); the 3x3 convolutions stay on cuDNN.
The single mid-block attention is one head of width C = 512 over N = h*w tokens: wider than the fused
tcgen05 kernel's TMEM budget (d_head <= 192), so it goes through vf_attn_fwd only when C <= 192 (reduced
test configurations) and through torch's scaled_dot_product_attention (a library call, off the hot path)
otherwise.  The Encoder (:368-459) + quant_conv + DiagonalGaussianDistribution (distributions.py:24-45) are
mirrored the same way for the encode -> decode round trip of the script's save loop
(scripts/VFace_inference_batch.py:456-459, :603-623); the training losses are out of scope.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import ops
from ...._autocast import no_autocast


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


def _nhwc(x):
    """(n, c, h, w) -> contiguous (n, h, w, c); free when x is already channels_last."""
    return x.permute(0, 2, 3, 1).contiguous()


def _gn_silu(norm, xt, silu=True):
    return ops.group_norm_nhwc(xt, norm.weight, norm.bias, norm.eps, norm.num_groups, silu=silu)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        x = ops.upsample_nearest2x(x)                    # F.interpolate(scale_factor=2.0, mode="nearest"), channels-last
        return self.conv(x) if self.with_conv else x


class Downsample(nn.Module):
    """Stride-2 3x3 convolution over the input padded by one zero row/column at the bottom/right (reference :60-79:
    `pad = (0, 1, 0, 1)`, "no asymmetric padding in torch conv")."""

    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)

    def forward(self, x):
        if not self.with_conv:
            return F.avg_pool2d(x, kernel_size=2, stride=2)
        return self.conv(F.pad(x, (0, 1, 0, 1), mode="constant", value=0))


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        if temb_channels:
            raise NotImplementedError("the first-stage decoder has no timestep embedding (temb_ch = 0)")
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.use_conv_shortcut = conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def forward(self, x, temb=None):
        xt = _nhwc(x)
        h = F.conv2d(_gn_silu(self.norm1, xt).permute(0, 3, 1, 2), self.conv1.weight, self.conv1.bias, padding=1)
        h = F.conv2d(_gn_silu(self.norm2, _nhwc(h)).permute(0, 3, 1, 2), self.conv2.weight, None, padding=1)
        bias = self.conv2.bias
        if self.in_channels == self.out_channels:
            skip = xt
        elif self.use_conv_shortcut:
            skip = _nhwc(self.conv_shortcut(x))
        else:
            skip = F.linear(xt, self.nin_shortcut.weight.reshape(self.out_channels, self.in_channels))
            bias = bias + self.nin_shortcut.bias
        return ops.add_bias(skip, _nhwc(h), bias).permute(0, 3, 1, 2)


class AttnBlock(nn.Module):
    """Single-head spatial self-attention of width C (reference :156-203)."""

    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, 1)
        self.k = nn.Conv2d(in_channels, in_channels, 1)
        self.v = nn.Conv2d(in_channels, in_channels, 1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, 1)

    def forward(self, x):
        n, c, hh, ww = x.shape
        xt = _nhwc(x)
        t = _gn_silu(self.norm, xt, silu=False).reshape(n, hh * ww, c)
        lin = lambda conv: F.linear(t, conv.weight.reshape(c, c), conv.bias)
        q, k, v = lin(self.q), lin(self.k), lin(self.v)
        # one head of width c on the repo's own attention kernel: the tcgen05 path takes c <= 192 and the wide-head form
        # 256 / 384 / 512 (REFace ddconfig: 512; csrc/vf_attn_stream.cu), the fp32 kernel c <= 256; anything else has no
        # kernel here and fails loudly inside ops.attention
        if x.dtype == torch.float32 and c > 256:
            # reference-precision (fp32) decode of a 512-wide head: scores in fp32 through torch (parity path only)
            o = F.scaled_dot_product_attention(q.unsqueeze(1), k.unsqueeze(1), v.unsqueeze(1), scale=float(c) ** -0.5).squeeze(1)
        else:
            o = ops.attention(q.contiguous(), k.contiguous(), v.contiguous(), heads=1, scale=float(c) ** -0.5)
        o = F.linear(o, self.proj_out.weight.reshape(c, c))
        return ops.add_bias(xt, o.reshape(n, hh, ww, c), self.proj_out.bias).permute(0, 3, 1, 2)


class Decoder(nn.Module):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 use_linear_attn=False, attn_type="vanilla", **ignorekwargs):
        super().__init__()
        if use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("only vanilla attention is used by the REFace first stage")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.give_pre_end, self.tanh_out = give_pre_end, tanh_out
        block_in = ch * ch_mult[self.num_resolutions - 1]
        curr_res = resolution // 2 ** (self.num_resolutions - 1)
        self.z_shape = (1, z_channels, curr_res, curr_res)
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            up = nn.Module()
            up.block, up.attn = block, attn
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    @no_autocast            # the script decodes under autocast("cuda") (VFace_inference_batch.py:407, :596): see _autocast.py
    def forward(self, z):
        z = z.to(self.conv_in.weight.dtype)
        if z.dtype == torch.float32:
            # reference-precision path: keep the cuDNN convolutions in true fp32 (no TF32)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._forward(z)
        return self._forward(z)

    def _forward(self, z):
        self.last_z_shape = z.shape
        if not getattr(self, "_weights_channels_last", False):
            self.to(memory_format=torch.channels_last)
            self._weights_channels_last = True
        h = self.conv_in(z.contiguous(memory_format=torch.channels_last))
        h = self.mid.block_1(h)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h)
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h)
                if len(self.up[i_level].attn) > 0:
                    h = self.up[i_level].attn[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        if self.give_pre_end:
            return h
        h = F.conv2d(_gn_silu(self.norm_out, _nhwc(h)).permute(0, 3, 1, 2), self.conv_out.weight, self.conv_out.bias, padding=1)
        return torch.tanh(h) if self.tanh_out else h


class Encoder(nn.Module):
    """Reference :368-459, same module tree (conv_in, down.{level}.block.{i}, down.{level}.downsample.conv,
    mid.block_1 / attn_1 / block_2, norm_out, conv_out): image (n, in_channels, H, W) -> moments (n, 2 z, H/8, W/8)."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, use_linear_attn=False,
                 attn_type="vanilla", **ignore_kwargs):
        super().__init__()
        if use_linear_attn or attn_type != "vanilla":
            raise NotImplementedError("only vanilla attention is used by the REFace first stage")
        self.ch, self.temb_ch = ch, 0
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.resolution, self.in_channels = resolution, in_channels
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)
        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block, attn = nn.ModuleList(), nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, temb_channels=0, dropout=dropout))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attn.append(AttnBlock(block_in))
            down = nn.Module()
            down.block, down.attn = block, attn
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, temb_channels=0, dropout=dropout)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, 2 * z_channels if double_z else z_channels, kernel_size=3, stride=1, padding=1)

    @no_autocast
    def forward(self, x):
        x = x.to(self.conv_in.weight.dtype)
        if x.dtype == torch.float32:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._forward(x)
        return self._forward(x)

    def _forward(self, x):
        if not getattr(self, "_weights_channels_last", False):
            self.to(memory_format=torch.channels_last)
            self._weights_channels_last = True
        h = self.conv_in(x.contiguous(memory_format=torch.channels_last))
        for i_level in range(self.num_resolutions):
            for i_block in range(self.num_res_blocks):
                h = self.down[i_level].block[i_block](h)
                if len(self.down[i_level].attn) > 0:
                    h = self.down[i_level].attn[i_block](h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_1(h)
        h = self.mid.attn_1(h)
        h = self.mid.block_2(h)
        return F.conv2d(_gn_silu(self.norm_out, _nhwc(h)).permute(0, 3, 1, 2), self.conv_out.weight, self.conv_out.bias, padding=1)


class DiagonalGaussianDistribution:
    """distributions.py:24-45: moments (n, 2 z, h, w) -> mean, logvar clamped to [-30, 20], std, var; sample / mode."""

    def __init__(self, parameters, deterministic=False):
        self.parameters = parameters
        self.mean, self.logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(self.logvar, -30.0, 20.0)
        self.deterministic = deterministic
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)
        if deterministic:
            self.var = self.std = torch.zeros_like(self.mean)

    def sample(self, generator=None):
        noise = torch.randn(self.mean.shape, generator=generator).to(device=self.parameters.device, dtype=self.mean.dtype)
        return self.mean + self.std * noise

    def mode(self):
        return self.mean


# configs/project_ffhq.yaml first_stage_config.params.ddconfig
REFACE_DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                       ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0)


class AutoencoderKLDecoder(nn.Module):
    """The decode half of AutoencoderKL (autoencoder.py:285-333): post_quant_conv + Decoder, same keys."""

    def __init__(self, ddconfig=None, embed_dim=4):
        super().__init__()
        cfg = dict(REFACE_DDCONFIG)
        if ddconfig:
            cfg.update(ddconfig)
        self.decoder = Decoder(**cfg)
        self.post_quant_conv = nn.Conv2d(embed_dim, cfg["z_channels"], 1)

    @no_autocast
    def decode(self, z):
        z = z.to(self.post_quant_conv.weight.dtype)
        if z.dtype == torch.float32:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self.decoder(self.post_quant_conv(z))
        return self.decoder(self.post_quant_conv(z))


class AutoencoderKL(AutoencoderKLDecoder):
    """Both halves of AutoencoderKL (autoencoder.py:285-333): encoder + quant_conv in front of the decode half, same
    state-dict keys as the reference's `first_stage_model.*` (minus the training loss)."""

    def __init__(self, ddconfig=None, embed_dim=4):
        super().__init__(ddconfig, embed_dim)
        cfg = dict(REFACE_DDCONFIG)
        if ddconfig:
            cfg.update(ddconfig)
        if not cfg.get("double_z", True):
            raise ValueError("AutoencoderKL needs double_z (autoencoder.py:303)")
        self.encoder = Encoder(**cfg)
        self.quant_conv = nn.Conv2d(2 * cfg["z_channels"], 2 * embed_dim, 1)

    @no_autocast
    def encode(self, x):
        x = x.to(self.quant_conv.weight.dtype)
        if x.dtype == torch.float32:
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))
        return DiagonalGaussianDistribution(self.quant_conv(self.encoder(x)))
