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


# ---- flow producer side: the 512^2 -> 64^2 contract ------------------------------------
# The shipped scripts estimate RAFT flow on the 512x512 frames and hand it to a hook that warps 64x64 feature
# maps, which raises inside warp_image; the line the authors left commented out
# (VFace_inference_batch.py:551-552, VFace_inference_single.py:773-775) resizes the VIDEO to H/f x W/f first, so
# that the estimator works at feature resolution and its displacements are already in feature pixels.  Both
# routes are provided; the hook consumes (B-1, 2, H/f, W/f) in feature-pixel units either way.

def resize_for_flow(video, f=8):
    """The authors' commented-out resize: frames (B, 3, H, W) -> (B, 3, H/f, W/f), bilinear, align_corners=False."""
    h, w = video.shape[-2] // f, video.shape[-1] // f
    return torch.nn.functional.interpolate(video, size=(h, w), mode="bilinear", align_corners=False)


def flow_to_feature_resolution(flow, size):
    """Flow estimated at image resolution -> feature resolution: (n, 2, H, W) image-pixel displacements (or the
    reference's list of (1, 2, H, W)) -> (n, 2, h, w) feature-pixel displacements.  The field is resampled
    bilinearly (align_corners=False, like the frames in resize_for_flow) and each component is scaled by the
    resolution ratio of ITS axis (channel 0 = x by w/W, channel 1 = y by h/H)."""
    if isinstance(flow, (list, tuple)):
        flow = torch.cat([f.reshape(1, 2, f.shape[-2], f.shape[-1]) for f in flow], dim=0)
    h, w = (size, size) if isinstance(size, int) else size
    H, W = flow.shape[-2:]
    if (H, W) == (h, w):
        return flow.float().contiguous()
    out = torch.nn.functional.interpolate(flow.float(), size=(h, w), mode="bilinear", align_corners=False)
    scale = torch.tensor([w / W, h / H], dtype=out.dtype, device=out.device).view(1, 2, 1, 1)
    return (out * scale).contiguous()


@torch.no_grad()
def return_flow(video, estimator=None, feature_size=None, num_flow_updates=20):
    """Mirror of return_flow (temporal_flow.py:163-188) with the estimator injected (the reference builds
    torchvision's raft_large at import time, :27, which needs downloaded weights): flow[i] maps frame i ->
    frame i+1, computed as estimator(frame[i+1], frame[i]) (:182, the reference's argument order) for ALL
    pairs in one batched call instead of the per-pair Python loop.  `estimator(img1, img2, num_flow_updates=)`
    returns the list of refinements (RAFT) or the flow itself.  With feature_size the result goes through
    flow_to_feature_resolution.  Returns a (B-1, 2, h, w) fp32 tensor, accepted everywhere a flow list is."""
    if estimator is None:
        raise RuntimeError("return_flow: pass the flow estimator (e.g. torchvision raft_large(...).eval()); "
                           "vface_b200 does not download weights")
    if video.shape[0] < 2:
        return torch.zeros(0, 2, video.shape[-2], video.shape[-1], device=video.device)
    try:
        out = estimator(video[1:], video[:-1], num_flow_updates=num_flow_updates)
    except TypeError:
        out = estimator(video[1:], video[:-1])
    flow = out[-1] if isinstance(out, (list, tuple)) else out
    return flow_to_feature_resolution(flow, feature_size) if feature_size is not None else flow.float().contiguous()
