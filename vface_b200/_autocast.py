"""torch.autocast and the mirrored modules.

scripts/VFace_inference_batch.py wraps sampling in `autocast("cuda")` by default (:400, :407-409; fp16).
Inside such a region every F.linear / F.conv2d re-casts its operands -- already-bf16 ones too -- and
returns float16, which the vface_b200 kernels (float32 / bfloat16 only, no CPU or eager fallback) refuse.
The mirrored modules therefore switch autocast off inside their forward and compute in their parameter
dtype, like the reference's own fp32-only islands (GroupNorm32, softmax) do; a UNet that still holds fp32
parameters when it is called under autocast takes the caller's request for 16-bit compute as bf16
(`adopt_autocast_dtype`), the 16-bit type of the tensor-core path.
"""
from __future__ import annotations

import functools
import warnings

import torch


def no_autocast(fn):
    """Run `fn` with CUDA autocast disabled (a no-op outside an autocast region)."""

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        if torch.is_autocast_enabled("cuda"):
            with torch.autocast("cuda", enabled=False):
                return fn(*args, **kwargs)
        return fn(*args, **kwargs)

    return wrapped


def autocast_requested() -> bool:
    return torch.is_autocast_enabled("cuda")


def adopt_autocast_dtype(module: torch.nn.Module, what: str) -> None:
    """Called by a mirrored network at the top of forward: fp32 parameters + an enclosing autocast region
    = the caller asked for 16-bit compute (the reference would run fp16 GEMMs/convolutions with fp32
    norms and softmax).  The network's parameters are cast to bfloat16 once, in place; state-dict keys
    are unchanged."""
    if not torch.is_autocast_enabled("cuda"):
        return
    p = next(module.parameters(), None)
    if p is None or p.dtype != torch.float32 or not p.is_cuda:
        return
    warnings.warn(f"vface_b200: {what} called under torch.autocast with float32 parameters -- casting it to bfloat16 "
                  "once (the kernels' 16-bit type); call .to(torch.bfloat16) yourself to silence this, or leave "
                  "autocast off (--precision full) for the float32 reference-precision path", stacklevel=3)
    module.to(torch.bfloat16)
