"""Minimal LatentDiffusion for standalone use of the hot path (bench, tests, synthetic clips).

In the reference the sampler is handed ldm.models.diffusion.ddpm.LatentDiffusion (a 2300-line
Lightning module with conditioning encoders, VAE and losses -- out of scope).
The sampler and hooks touch only: .model.diffusion_model, .apply_model (ddpm.py:1519 -> :1609 ->
DiffusionWrapper.forward :2245-2247, a pass-through on the crossattn path), .num_timesteps,
.parameterization and the schedule buffers betas / alphas_cumprod / alphas_cumprod_prev
(ddpm.py:255-277).  This class provides exactly that around a vface_b200 UNetModel.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from .ldm.modules.diffusionmodules.openaimodel import UNetModel
from .ldm.modules.diffusionmodules.util import make_beta_schedule

# models/REFace/configs/project_ffhq.yaml:33-55
REFACE_UNET_CONFIG = dict(image_size=32, in_channels=9, out_channels=4, model_channels=320,
                          attention_resolutions=[4, 2, 1], num_res_blocks=2, channel_mult=[1, 2, 4, 4],
                          num_heads=8, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                          use_checkpoint=True, legacy=False)


class DiffusionWrapper(nn.Module):
    def __init__(self, unet, conditioning_key="crossattn"):
        super().__init__()
        self.diffusion_model = unet
        self.conditioning_key = conditioning_key

    def forward(self, x, t, c_concat=None, c_crossattn=None, return_features=False):
        if self.conditioning_key != "crossattn":
            raise NotImplementedError("only the crossattn conditioning path is used by VFace")
        cc = c_crossattn[0] if len(c_crossattn) == 1 else torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, t, context=cc, return_features=return_features)


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config=None, timesteps=1000, linear_start=0.00085, linear_end=0.012, unet=None,
                 first_stage_config=None, first_stage_encoder=False):
        super().__init__()
        cfg = dict(REFACE_UNET_CONFIG)
        if unet_config:
            cfg.update(unet_config)
        self.model = DiffusionWrapper(unet if unet is not None else UNetModel(**cfg))
        self.num_timesteps = timesteps
        self.parameterization = "eps"
        self.scale_factor = 0.18215
        # the step after the path: only built on request, `{}` = the REFace ddconfig
        if first_stage_config is not None:
            from .ldm.modules.diffusionmodules.model import AutoencoderKL, AutoencoderKLDecoder
            cls = AutoencoderKL if first_stage_encoder else AutoencoderKLDecoder
            self.first_stage_model = cls(first_stage_config or None)
        betas = make_beta_schedule("linear", timesteps, linear_start=linear_start, linear_end=linear_end)
        alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
        f32 = lambda a: torch.tensor(a, dtype=torch.float32)
        self.register_buffer("betas", f32(betas))
        self.register_buffer("alphas_cumprod", f32(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", f32(alphas_cumprod_prev))

    @property
    def device(self):
        return self.betas.device

    def apply_model(self, x_noisy, t, cond, return_ids=False, return_features=False):
        if isinstance(cond, dict):
            return self.model(x_noisy, t, **cond)
        if not isinstance(cond, list):
            cond = [cond]
        return self.model(x_noisy, t, c_crossattn=cond, return_features=return_features)

    def q_sample(self, x_start, t, noise=None):
        noise = torch.randn_like(x_start) if noise is None else noise
        a = self.alphas_cumprod[t].sqrt().view(-1, 1, 1, 1)
        s = (1.0 - self.alphas_cumprod[t]).sqrt().view(-1, 1, 1, 1)
        return a * x_start + s * noise

    def decode_first_stage(self, z):
        """ddpm.py:1277-1284 (plain path) -> AutoencoderKL.decode: x = decoder(post_quant_conv(z / scale_factor))."""
        if not hasattr(self, "first_stage_model"):
            raise RuntimeError("LatentDiffusion was built without first_stage_config")
        p = next(self.first_stage_model.parameters())
        return self.first_stage_model.decode((1.0 / self.scale_factor * z).to(p.dtype))

    def encode_first_stage(self, x):
        """ddpm.py:1307-1330 (plain path) -> AutoencoderKL.encode: the posterior over the latents of image x."""
        fs = getattr(self, "first_stage_model", None)
        if fs is None or not hasattr(fs, "encode"):
            raise RuntimeError("LatentDiffusion was built without first_stage_encoder=True")
        p = next(fs.parameters())
        return fs.encode(x.to(p.dtype))

    def get_first_stage_encoding(self, encoder_posterior):
        """ddpm.py:782-791: scale_factor * (a sample of the posterior, or the tensor itself)."""
        z = encoder_posterior.sample() if hasattr(encoder_posterior, "sample") else encoder_posterior
        return self.scale_factor * z

    def to_compute_dtype(self, dtype):
        """Cast the UNet parameters (bf16 on the throughput path); schedule buffers stay fp32."""
        self.model.diffusion_model.to(dtype)
        return self
