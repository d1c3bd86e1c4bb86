"""The REFace / Paint-by-Example denoising UNet on the vface_b200 kernels.

Host-side mirror of REFace/ldm/modules/diffusionmodules/openaimodel.py: UNetModel :528-907,
ResBlock :163-275, Upsample :91-119, Downsample :134-160, TimestepEmbedSequential :74-88.
Constructor keywords are the reference's (models/REFace/configs/project_ffhq.yaml:33-55) and the
module tree -- hence every state-dict key -- is identical, so the reference checkpoint loads as is.

What differs is the execution plan: activations are channels_last (NHWC) in the parameter dtype
(bf16 on the throughput path), so the 16 SpatialTransformers see their (b, hw, c) token view
without a copy, attention never materialises N x N, and gradient checkpointing (a training device,
`use_checkpoint`) is a no-op.  Convolutions, GroupNorm and the projection GEMMs stay on
cuDNN/cuBLAS (outside the four hand-written subsystems).

Options of the reference constructor that the VFace configuration never enables raise
NotImplementedError instead of being silently ignored.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .... import ops
from ...._autocast import adopt_autocast_dtype, no_autocast
from ..attention import SpatialTransformer
from .util import normalization, timestep_embedding, zero_module


class TimestepBlock(nn.Module):
    """Marker: forward(x, emb)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None):
        """x may be a (h, skip) pair: the decoder's channel concatenation, consumed in place by the
        leading ResBlock (every output block starts with one)."""
        for layer in self:
            if isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context)
            elif isinstance(layer, nn.Conv2d):
                x = ops.conv2d_nhwc(x, layer)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if use_conv:
            self.conv = nn.Conv2d(self.channels, self.out_channels, 3, padding=padding)

    def forward(self, x):
        x = ops.upsample_nearest2x(x)                    # F.interpolate(scale_factor=2, mode="nearest"), channels-last
        return ops.conv2d_nhwc(x, self.conv) if self.use_conv else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.dims = dims
        if use_conv:
            self.op = nn.Conv2d(self.channels, self.out_channels, 3, stride=2, padding=padding)
        else:
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        return ops.conv2d_nhwc(x, self.op) if self.use_conv else self.op(x)


class ResBlock(TimestepBlock):
    """GN-SiLU-conv3x3, + emb projection, GN-SiLU-(dropout)-conv3x3 (zero-init), + skip."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        if up or down or use_scale_shift_norm or dims != 2:
            raise NotImplementedError("resblock_updown / scale-shift norm / non-2D are not used by the VFace configuration")
        self.channels = channels
        self.emb_channels = emb_channels
        self.dropout = dropout
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        self.use_checkpoint = use_checkpoint
        self.use_scale_shift_norm = use_scale_shift_norm
        self.updown = False
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       nn.Conv2d(channels, self.out_channels, 3, padding=1))
        self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(nn.Conv2d(self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = nn.Conv2d(channels, self.out_channels, 1)

    def forward(self, x, emb):
        return self._forward(x, emb)

    def _split_skip_weight(self, c1):
        """The 1x1 skip convolution's weight as two contiguous (out, c1) / (out, c - c1) matrices (the two sources of a
        never-materialised channel concatenation), cached until the parameter changes."""
        wp = self.skip_connection.weight
        key = (wp.data_ptr(), wp._version, wp.dtype, wp.device, c1)
        if getattr(self, "_skip_split_key", None) != key:
            w = wp.detach().reshape(self.out_channels, self.channels)
            self._skip_split = (w[:, :c1].contiguous(), w[:, c1:].contiguous())
            self._skip_split_key = key
        return self._skip_split

    @staticmethod
    def _bias_sum(owner, name, a, b):
        """a + b of two bias parameters, cached on `owner` until either changes (one tiny add kernel per block and step
        otherwise: a 32-frame step launched ~60 of them)."""
        key = (a.data_ptr(), a._version, b.data_ptr(), b._version, a.dtype, a.device)
        hit = owner.__dict__.get(name)
        if hit is None or hit[0] != key:
            hit = (key, (a.detach() + b.detach()).contiguous())
            owner.__dict__[name] = hit
        return hit[1]

    def _emb_plus_conv_bias(self, emb, conv1):
        """emb_layers(emb) + conv1.bias (reference :265-273: `h = in_layers(x)`'s bias and `h + emb_out`): the two biases
        are summed once, and SiLU(emb) -- the same tensor for every ResBlock of a forward -- is taken from the cache
        UNetModel.forward leaves on `emb`."""
        act, lin = self.emb_layers[0], self.emb_layers[1]
        if type(act) is not nn.SiLU or type(lin) is not nn.Linear or lin.bias is None or conv1.bias is None:
            return self.emb_layers(emb) + conv1.bias
        se = getattr(emb, "_vf_silu", None)
        if se is None:
            se = F.silu(emb)
        return F.linear(se, lin.weight, self._bias_sum(self, "_emb_conv_bias", lin.bias, conv1.bias))

    def _forward(self, x, emb):
        """GN-SiLU-conv, + emb, GN-SiLU-conv, + skip (reference :255-275) on channels-last activations:
        the two GroupNorm+SiLU are one fused kernel each, the `+ emb_out` and the first conv bias ride
        inside the second GroupNorm, and the second conv bias rides inside the residual add."""
        x2t = None
        if isinstance(x, tuple):
            # decoder skip concatenation th.cat([h, hs.pop()], dim=1) (reference :899), never materialised:
            # the first GroupNorm reads both sources, the 1x1 skip convolution becomes two accumulating GEMMs
            x, x2 = x
            x2t = x2.permute(0, 2, 3, 1).contiguous()
        n, c, hh, ww = x.shape
        xt = x.permute(0, 2, 3, 1).contiguous()                  # NHWC; a view for channels_last input
        gn1, conv1 = self.in_layers[0], self.in_layers[2]
        gn2, conv2 = self.out_layers[0], self.out_layers[3]
        g1 = ops.group_norm_nhwc(xt, gn1.weight, gn1.bias, gn1.eps, gn1.num_groups, silu=True, x2=x2t)
        h1 = F.conv2d(g1.permute(0, 3, 1, 2), conv1.weight, None, padding=1)
        add = self._emb_plus_conv_bias(emb, conv1).type(h1.dtype)
        g2 = ops.group_norm_nhwc(h1.permute(0, 2, 3, 1).contiguous(), gn2.weight, gn2.bias, gn2.eps, gn2.num_groups,
                                 silu=True, add_nc=add)
        h2 = F.conv2d(g2.permute(0, 3, 1, 2), conv2.weight, None, padding=1).permute(0, 2, 3, 1).contiguous()
        bias = conv2.bias
        fuse = h2.dtype == torch.bfloat16       # bf16: 1x1 skip convolution + biases + residual in the library GEMM's epilogue
        rows = n * hh * ww
        if x2t is not None:
            if isinstance(self.skip_connection, nn.Identity) or self.skip_connection.kernel_size != (1, 1):
                raise NotImplementedError("a concatenated ResBlock input needs the 1x1 skip convolution")
            bias = self._bias_sum(self, "_out_skip_bias", bias, self.skip_connection.bias)
            w1, w2 = self._split_skip_weight(c)
            if fuse:
                out = ops.linear_residual(xt.reshape(rows, c), w1, bias, h2.reshape(rows, self.out_channels))
                out = ops.linear_residual(x2t.reshape(rows, self.channels - c), w2, None, out)
                return out.reshape(n, hh, ww, self.out_channels).permute(0, 3, 1, 2)
            skip = torch.mm(xt.reshape(rows, c), w1.t())
            skip.addmm_(x2t.reshape(rows, self.channels - c), w2.t())
            skip = skip.reshape(n, hh, ww, self.out_channels)
        elif isinstance(self.skip_connection, nn.Identity):
            skip = xt
        elif self.skip_connection.kernel_size == (1, 1):
            bias = self._bias_sum(self, "_out_skip_bias", bias, self.skip_connection.bias)
            w = self.skip_connection.weight.reshape(self.out_channels, c)
            if fuse and w.is_contiguous():
                out = ops.linear_residual(xt.reshape(rows, c), w, bias, h2.reshape(rows, self.out_channels))
                return out.reshape(n, hh, ww, self.out_channels).permute(0, 3, 1, 2)
            skip = F.linear(xt, w)
        else:
            skip = self.skip_connection(x).permute(0, 2, 3, 1).contiguous()
        out = ops.add_bias(skip, h2, bias)
        return out.permute(0, 3, 1, 2)


class UNetModel(nn.Module):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks,
                 attention_resolutions, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1,
                 context_dim=None, n_embed=None, legacy=True, add_conv_in_front_of_unet=False,
                 sep_head_att=False, land_mark_id_seperate_layers=False, head_splits=None):
        super().__init__()
        unsupported = dict(dims=dims != 2, num_classes=num_classes is not None, n_embed=n_embed is not None,
                           use_scale_shift_norm=use_scale_shift_norm, resblock_updown=resblock_updown,
                           add_conv_in_front_of_unet=add_conv_in_front_of_unet, sep_head_att=sep_head_att,
                           land_mark_id_seperate_layers=land_mark_id_seperate_layers,
                           no_spatial_transformer=not use_spatial_transformer)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"UNetModel options outside the VFace configuration: {bad}")
        if context_dim is None:
            raise ValueError("use_spatial_transformer needs context_dim")
        if num_heads == -1 and num_head_channels == -1:
            raise ValueError("either num_heads or num_head_channels has to be set")
        if not isinstance(context_dim, int):
            context_dim = list(context_dim)   # OmegaConf ListConfig
            raise NotImplementedError("list-valued context_dim is not used by the VFace configuration")
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads

        self.image_size = image_size
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = list(attention_resolutions)
        self.dropout = dropout
        self.channel_mult = list(channel_mult)
        self.conv_resample = conv_resample
        self.num_classes = None
        self.use_checkpoint = use_checkpoint
        self.num_heads = num_heads
        self.num_head_channels = num_head_channels
        self.num_heads_upsample = num_heads_upsample
        self.predict_codebook_ids = False

        def heads_for(ch):
            if num_head_channels == -1:
                return num_heads, ch // num_heads
            return ch // num_head_channels, num_head_channels

        def transformer(ch):
            nh, dh = heads_for(ch)
            return SpatialTransformer(ch, nh, dh, depth=transformer_depth, context_dim=context_dim)

        def res(cin, cout):
            return ResBlock(cin, time_embed_dim, dropout, out_channels=cout, use_checkpoint=use_checkpoint)

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(nn.Linear(model_channels, time_embed_dim), nn.SiLU(),
                                        nn.Linear(time_embed_dim, time_embed_dim))

        # ---- encoder ----------------------------------------------------------------------------
        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(nn.Conv2d(in_channels, model_channels, 3, padding=1))])
        skip_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    layers.append(transformer(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2

        # ---- bottleneck -------------------------------------------------------------------------
        self.middle_block = TimestepEmbedSequential(res(ch, ch), transformer(ch), res(ch, ch))

        # ---- decoder ----------------------------------------------------------------------------
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in self.attention_resolutions:
                    layers.append(transformer(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))

        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(nn.Conv2d(model_channels, out_channels, 3, padding=1)))

    @property
    def dtype(self):
        return self.input_blocks[0][0].weight.dtype

    def forward(self, x, timesteps=None, context=None, y=None, return_features=False, **kwargs):
        """(N, in_channels, H, W), (N,) timesteps, (N, M, context_dim) -> (N, out_channels, H, W)."""
        if y is not None:
            raise NotImplementedError("class-conditional UNet is not used by the VFace configuration")
        # scripts/VFace_inference_batch.py:400-409 samples under autocast("cuda") by default: fp32 parameters become
        # bf16 once (the caller asked for 16-bit compute), and autocast itself stays off inside -- it would turn every
        # F.linear / F.conv2d result into float16, which the kernels do not take (vface_b200/_autocast.py)
        adopt_autocast_dtype(self, "UNetModel.forward")
        return self._forward_no_autocast(x, timesteps, context, return_features)

    @no_autocast
    def _forward_no_autocast(self, x, timesteps, context, return_features):
        dt = self.dtype
        if dt == torch.float32:
            # reference-precision path: keep cuDNN convolutions in true fp32 (no TF32)
            with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
                return self._forward(x, timesteps, context, return_features)
        return self._forward(x, timesteps, context, return_features)

    def _forward(self, x, timesteps, context, return_features):
        dt = self.dtype
        if not getattr(self, "_weights_channels_last", False):
            self.to(memory_format=torch.channels_last)       # conv weights NHWC once; keys/values unchanged
            self._weights_channels_last = True
        emb = self.time_embed(timestep_embedding(timesteps, self.model_channels).to(dt))
        emb._vf_silu = F.silu(emb)              # every ResBlock's emb_layers starts with the same SiLU(emb): once, not 22 times
        context = context.to(dt)
        h = x.to(dt).contiguous(memory_format=torch.channels_last)
        hs = []
        for module in self.input_blocks:
            h = module(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        features = []
        for module in self.output_blocks:
            h = module((h, hs.pop()), emb, context)
            if return_features:
                features.append(h)
        gn = self.out[0]
        ht = ops.group_norm_nhwc(h.permute(0, 2, 3, 1).contiguous(), gn.weight, gn.bias, gn.eps, gn.num_groups, silu=True)
        conv = self.out[2]
        if dt == torch.bfloat16 and conv.out_channels == 4 and conv.in_channels % 32 == 0 and x.dtype == torch.float32:
            out = ops.conv3x3_out_f32(ht, conv)              # eps in fp32: CFG multiplies a bf16 rounding by 3.6
        else:
            out = ops.conv2d_nhwc(ht.permute(0, 3, 1, 2), conv).to(x.dtype).contiguous()
        return (out, features) if return_features else out
