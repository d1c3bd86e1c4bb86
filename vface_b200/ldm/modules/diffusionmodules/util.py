"""Schedule maths and small layers of the diffusion UNet.

Mirror of the pieces of REFace/ldm/modules/diffusionmodules/util.py that the hot path uses
(SURVEY.md section 8 rows a1, a6, a13): make_beta_schedule :21-25, make_ddim_timesteps :46-60,
make_ddim_sampling_parameters :63-74, timestep_embedding :151-171, GroupNorm32 :214-216,
noise_like :264-267.  Same names and argument meaning; written from the formulas, not the code.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """fp64 betas; only the 'linear' (sqrt-space linear) schedule of project_ffhq.yaml is on the path."""
    if schedule != "linear":
        raise NotImplementedError(f"beta schedule '{schedule}' is outside the VFace hot path")
    lo, hi = linear_start ** 0.5, linear_end ** 0.5
    return (torch.linspace(lo, hi, n_timestep, dtype=torch.float64) ** 2).numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    """Uniform DDIM sub-sequence, shifted by +1 (reference util.py:46-60)."""
    if ddim_discr_method != "uniform":
        raise NotImplementedError(f"ddim discretisation '{ddim_discr_method}' is outside the VFace hot path")
    stride = num_ddpm_timesteps // num_ddim_timesteps
    steps = np.arange(0, num_ddpm_timesteps, stride, dtype=np.int64) + 1
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps}")
    return steps


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    """sigmas, alphas, alphas_prev for the selected steps.

    `alphacums` is the model's fp32 alphas_cumprod; alphas keeps that fp32 precision, alphas_prev
    is the same values shifted by one (first entry alphacums[0]) held in fp64, and sigma follows
    eta * sqrt((1-a_prev)/(1-a) * (1 - a/a_prev))  (reference util.py:63-74).
    """
    acp = torch.as_tensor(alphacums).detach().cpu()
    alphas = acp[torch.as_tensor(ddim_timesteps, dtype=torch.long)]
    alphas_prev = np.asarray([acp[0].item()] + acp[torch.as_tensor(ddim_timesteps[:-1], dtype=torch.long)].tolist())
    a64 = alphas.double().numpy()
    # The reference evaluates this with a fp32 tensor (alphas) and a fp64 array (alphas_prev); numpy
    # defers `array / tensor` to Tensor.__rtruediv__ = reciprocal(tensor) * array, so 1/(1 - alphas)
    # is rounded to fp32 and every other term is fp64.  Reproduced so sigma_t is bit-identical (eta > 0).
    recip = (1. - alphas).reciprocal().double().numpy()
    sigmas = eta * np.sqrt((recip * (1 - alphas_prev)) * (1 - a64 / alphas_prev))
    if verbose:
        print(f"Selected alphas for ddim sampler: a_t: {alphas}; a_(t-1): {alphas_prev}")
        print(f"For the chosen value of eta, which is {eta}, this results in the following sigma_t schedule "
              f"for ddim sampler {sigmas}")
    return sigmas, alphas, alphas_prev


_FREQS = {}


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """Sinusoidal embedding [cos | sin] of (N,) timesteps -> (N, dim) fp32 (reference util.py:151-171)."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    # computed on the host like the reference (bit-identical frequencies), but only once per (half, period, device): the
    # per-call host -> device copy of the reference is also what makes a step impossible to capture in a CUDA graph
    key = (half, max_period, timesteps.device)
    freqs = _FREQS.get(key)
    if freqs is None:
        freqs = _FREQS[key] = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(timesteps.device)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = F.pad(emb, (0, 1))
    return emb


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    """GroupNorm with fp32 statistics; output in the input dtype (reference util.py:214-216)."""

    def forward(self, x):
        if x.dtype == torch.float32:
            return super().forward(x)
        # bf16 activations: torch's kernel accumulates in fp32 and rounds once at the output,
        # which is what x.float() -> GN -> .type(x.dtype) does, without the extra round trip.
        return F.group_norm(x, self.num_groups, self.weight, self.bias, self.eps)


def normalization(channels):
    return GroupNorm32(32, channels)


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)
