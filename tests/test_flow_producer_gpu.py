"""SURVEY.md 8(f) row 3 on the GPU: the flow producer in front of the hot path.

The reference estimates flow with torchvision's RAFT-large, one frame pair per call in a Python loop
(scripts/temporal_flow.py:27-37, :163-188), on 512 x 512 frames, and hands the result to a hook that warps 64 x 64 feature
maps (the shape bug of SURVEY.md F5).  vface_b200.scripts.temporal_flow.return_flow runs the SAME estimator once over all
pairs and brings the field to feature resolution / feature-pixel units.  Weights are not available offline, so the
estimator here is the RAFT-large ARCHITECTURE with random weights: what is pinned is the batching contract (batched ==
the reference's per-pair loop, same argument order), the resolution contract, and that the result drives the warp hook."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _raft():
    from torchvision.models.optical_flow import raft_large
    torch.manual_seed(0)
    return raft_large(weights=None, progress=False).cuda().eval()


def test_batched_raft_equals_reference_loop_and_feeds_the_hook():
    from vface_b200 import ops
    from vface_b200.scripts import temporal_flow as tf
    model = _raft()
    g = torch.Generator().manual_seed(4)
    video = (torch.rand(4, 3, 256, 256, generator=g) * 2 - 1).cuda()       # frames in [-1, 1] as the dataset yields them
    updates = 2
    # the reference loop (temporal_flow.py:170-186): flow_i = model(frame[i+1], frame[i])[-1], one pair per call
    with torch.no_grad():
        loop = [model(video[i + 1:i + 2], video[i:i + 1], num_flow_updates=updates)[-1] for i in range(video.shape[0] - 1)]
    want = torch.cat(loop, dim=0)
    got = tf.return_flow(video, estimator=model, num_flow_updates=updates)
    assert tuple(got.shape) == (3, 2, 256, 256) and got.dtype == torch.float32
    # same network, same pairs, batch 3 instead of 1: cuDNN may pick other kernels, so fp32 round-off only
    assert (got - want).abs().max().item() <= 2e-3 * max(1.0, want.abs().max().item())
    # resolution contract: image-pixel flow at 256^2 -> feature-pixel flow at 32^2 (factor 8), both routes
    feat = tf.return_flow(video, estimator=model, feature_size=32, num_flow_updates=updates)
    assert tuple(feat.shape) == (3, 2, 32, 32)
    assert torch.allclose(feat, tf.flow_to_feature_resolution(want, 32), atol=2e-3 * max(1.0, want.abs().max().item()))
    small = tf.resize_for_flow(video, f=8)
    assert tuple(small.shape) == (4, 3, 32, 32)
    # and the field drives the warp hook on feature maps of that resolution (the reference's own call raises here, F5)
    x = torch.randn(4, 32 * 32, 64, device="cuda").bfloat16()
    out = ops.flow_warp_blend(x, feat, 0.8, 32, 32)
    assert tuple(out.shape) == tuple(x.shape) and torch.isfinite(out.float()).all()
    assert torch.equal(out[0], x[0])


def test_return_flow_needs_an_estimator_and_handles_one_frame():
    from vface_b200.scripts import temporal_flow as tf
    with pytest.raises(RuntimeError):
        tf.return_flow(torch.zeros(2, 3, 64, 64, device="cuda"))
    out = tf.return_flow(torch.zeros(1, 3, 64, 64, device="cuda"), estimator=lambda a, b: a[:, :2])
    assert tuple(out.shape) == (0, 2, 64, 64)
