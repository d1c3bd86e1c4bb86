"""FSAI entry point with the reference's name and argument meaning.

Mirror of combine_fft_high_low in REFace/scripts/face_swap_utils.py:425-464; the other helpers of that
file (plots, unused fusion variants) are outside the hot path.
"""
from __future__ import annotations

import torch

from .. import ops


def combine_fft_high_low(q1: torch.Tensor, q2: torch.Tensor, split_ratio: float = 0.5) -> torch.Tensor:
    """High-frequency bins [int(d*split_ratio), d) of q1 (the donor), low bins of q2, along the last
    axis; returns a new (b, n, d) tensor.  The reference casts to fp32 and returns fp32
    (face_swap_utils.py:443-462); here fp32 inputs give fp32, bf16 inputs are filtered in fp32
    registers and rounded once to bf16."""
    if q1.dim() == 2:
        return ops.fsai_blend(q1[None], q2[None], split_ratio)[0]
    return ops.fsai_blend(q1, q2, split_ratio)
