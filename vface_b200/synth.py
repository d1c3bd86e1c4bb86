"""Deterministic synthetic weights and clips (no checkpoints or datasets are available offline).

Weights: random-init of the named architecture.  A freshly constructed UNet is vacuous --
zero_module zeroes UNet.out[-1], every ResBlock.out_layers[-1] and every SpatialTransformer.proj_out,
so the network outputs exactly 0 and attention is disconnected.  `synth_state_dict`
therefore re-draws EVERY tensor from a recipe keyed by (seed, parameter name, shape) only, so the
reference UNet, the oracle port and the vface_b200 UNet get bit-identical weights regardless of
construction order, and residual branches are scaled so that ||eps|| ~ 1 and the net is not chaotic.

Clips: the synthetic recipe of SURVEY.md 8(d) config 1 (shapes of the named configuration).
"""
from __future__ import annotations

import math
import zlib
from typing import Dict, Tuple

import numpy as np
import torch


def _gen(seed: int, name: str) -> torch.Generator:
    return torch.Generator(device="cpu").manual_seed((seed * 1_000_003 + zlib.crc32(name.encode())) % (2 ** 63))


def synth_tensor(seed: int, name: str, shape: Tuple[int, ...]) -> torch.Tensor:
    g = _gen(seed, name)
    is_norm = (".norm" in name or "in_layers.0." in name or "out_layers.0." in name or name.startswith("out.0.")
               or name.startswith("norm"))
    if name.endswith(".bias"):
        return torch.randn(shape, generator=g) * (0.1 if is_norm else 0.02)
    if len(shape) == 1:   # norm gains
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    fan_in = int(np.prod(shape[1:]))
    gain = 1.0
    # residual-branch outputs (zero-initialised in the reference): keep them a perturbation
    if name.endswith("out_layers.3.weight") or name.endswith("proj_out.weight") or name.endswith("to_out.0.weight") \
            or name.endswith("ff.net.2.weight"):
        gain = 0.5
    return torch.randn(shape, generator=g) * (gain / math.sqrt(fan_in))


def synth_state_dict(reference_state: Dict[str, torch.Tensor], seed: int = 1) -> Dict[str, torch.Tensor]:
    """New fp32 state dict with the keys/shapes of `reference_state`."""
    return {k: synth_tensor(seed, k, tuple(v.shape)) for k, v in reference_state.items()}


def synth_clip(frames: int, seed: int = 7, steps=None, hw: int = 64, flow_kind: str = "smooth"):
    """Inputs of DDIMSampler.sample for a `frames`-frame 512x512 clip (latents hw x hw):
    x_T, inpaint_image ~ N(0,1); inpaint_mask in {0,1}; c, target_cond ~ N(0,1) (F,1,768); uc one
    N(0,1) vector repeated; flow: F-1 fields (1,2,hw,hw) in feature-pixel units; inversion latents
    N(0,1) per timestep (seed 11+t) for the timesteps in `steps`."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    clip = dict(
        x_T=r(frames, 4, hw, hw),
        inpaint_image=r(frames, 4, hw, hw),
        inpaint_mask=(torch.rand(frames, 1, hw, hw, generator=g) > 0.5).float(),
        c=r(frames, 1, 768),
        target_cond=r(frames, 1, 768),
        uc=r(1, 1, 768).repeat(frames, 1, 1),
    )
    gf = torch.Generator(device="cpu").manual_seed(seed + 6)
    ys, xs = torch.meshgrid(torch.arange(hw, dtype=torch.float32), torch.arange(hw, dtype=torch.float32), indexing="ij")
    flows = []
    for i in range(frames - 1):
        if flow_kind == "integer":
            f = torch.randint(-4, 5, (1, 2, hw, hw), generator=gf).float()
        else:
            fx = 3 * torch.sin(2 * math.pi * ys / hw + 0.3 * i) + 0.5 * torch.randn(hw, hw, generator=gf)
            fy = 3 * torch.cos(2 * math.pi * xs / hw - 0.2 * i) + 0.5 * torch.randn(hw, hw, generator=gf)
            f = torch.stack([fx, fy])[None]
        flows.append(f)
    clip["flow"] = flows
    if steps is not None:
        inv = {}
        for t in steps:
            gt = torch.Generator(device="cpu").manual_seed(11 + int(t))
            inv[int(t)] = torch.randn(frames, 4, hw, hw, generator=gt)
        clip["inversion"] = inv
    return clip
