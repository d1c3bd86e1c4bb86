"""Flow-guided alignment entry points with the reference's names.

Mirror of warp_image / align_by_flow in REFace/scripts/temporal_flow.py:40-53, :222-237.  RAFT flow
estimation (return_flow, :163-188) is a pre-step outside the hot path: flow is
an input here, at FEATURE resolution and in feature-pixel units.
"""
from __future__ import annotations

import torch

from .. import ops


def _to_tokens(x):
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, h * w, c).contiguous(), h, w


def _to_image(t, h, w):
    b, n, c = t.shape
    return t.reshape(b, h, w, c).permute(0, 3, 1, 2)


@torch.no_grad()
def align_by_flow(x_prev=None, flow=None, alpha=0.5):
    """x_prev (B, C, H, W); flow: list of B-1 tensors (1, 2, H, W) or a (B-1, 2, H, W) tensor.
    out[0] = x[0]; out[i+1] = alpha*x[i+1] + (1-alpha)*warp(x[i], flow[i])."""
    tok, h, w = _to_tokens(x_prev)
    return _to_image(ops.flow_warp_blend(tok, flow, alpha, h, w), h, w)


@torch.no_grad()
def warp_image(img, flow):
    """img (B, C, H, W), flow (B, 2, H, W): bilinear, border-padded sample of img at p + flow(p).
    Expressed with the blend kernel: frame pair (img[i], .) with alpha = 0."""
    outs = []
    for i in range(img.shape[0]):
        tok, h, w = _to_tokens(img[i:i + 1])
        pair = torch.cat([tok, torch.zeros_like(tok)], dim=0)
        outs.append(ops.flow_warp_blend(pair, flow[i:i + 1], 0.0, h, w)[1:2])
    return _to_image(torch.cat(outs, dim=0), img.shape[2], img.shape[3])
