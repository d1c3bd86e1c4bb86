"""The fused memory-bound glue kernels (vf_norm.cu) against plain PyTorch fp32 references of the same ops:
GroupNorm32 (+SiLU, + per-sample additive vector, + never-materialised channel concatenation), residual-add +
LayerNorm (all four forms, ragged row counts for the multi-row-per-warp kernels), GEGLU, add+bias.  The reference
operations are the ones the reference UNet executes (openaimodel.py:201-205, :236-239, :265-275; util.py:214-216;
attention.py:37-45, :239-243, :278-288); tolerances are stated per test."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


def _dev():
    return torch.device("cuda:0")


def _mk(shape, seed, dtype, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(_dev())


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("n,hw,c", [(3, 4096, 320), (2, 1024, 640), (5, 64, 1280), (2, 100, 96)])
@pytest.mark.parametrize("silu,with_add", [(True, False), (True, True), (False, False)])
def test_group_norm_nhwc(dtype, tol, n, hw, c, silu, with_add):
    from vface_b200 import ops
    x = _mk((n, hw, c), 1 + hw + c, dtype, 1.5) + 0.3
    w, b = _mk((c,), 2, dtype), _mk((c,), 3, dtype)
    add = _mk((n, c), 4, dtype) if with_add else None
    got = ops.group_norm_nhwc(x, w, b, 1e-5, 32, silu=silu, add_nc=add)
    xf = x.float() + (add.float()[:, None, :] if add is not None else 0.0)
    want = F.group_norm(xf.permute(0, 2, 1), 32, w.float(), b.float(), 1e-5).permute(0, 2, 1)
    if silu:
        want = F.silu(want)
    assert got.shape == x.shape and got.dtype == dtype
    assert (got.float() - want).abs().max().item() < tol * max(1.0, want.abs().max().item() / 2.0)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("c1,c2", [(320, 320), (640, 320), (1280, 640)])
def test_group_norm_over_concatenation(dtype, tol, c1, c2):
    """GroupNorm of th.cat([h, skip], 1) without materialising the concatenation -- including 640 + 320, where the
    group boundaries (30 channels) do not align with the boundary between the two sources."""
    from vface_b200 import ops
    n, hw = 2, 256
    x1, x2 = _mk((n, hw, c1), 5, dtype), _mk((n, hw, c2), 6, dtype, 2.0)
    c = c1 + c2
    w, b = _mk((c,), 7, dtype), _mk((c,), 8, dtype)
    got = ops.group_norm_nhwc(x1, w, b, 1e-5, 32, silu=True, x2=x2)
    cat = torch.cat([x1.float(), x2.float()], dim=-1)
    want = F.silu(F.group_norm(cat.permute(0, 2, 1), 32, w.float(), b.float(), 1e-5).permute(0, 2, 1))
    assert tuple(got.shape) == (n, hw, c)
    assert (got.float() - want).abs().max().item() < tol * max(1.0, want.abs().max().item() / 2.0)


def test_group_norm_full_size_properties():
    """Size-independent properties at the shape of a 32-frame step (96 x 64 x 64 x 320): the output does not change under a
    positive rescaling plus shift of a whole sample (eps aside), every (sample, group) of the normalised tensor has mean
    beta-weighted 0 and variance 1 when gamma = 1, beta = 0, and the per-sample additive vector equals adding it first."""
    from vface_b200 import ops
    dtype = torch.float32
    n, hw, c = 96, 4096, 320
    x = _mk((n, hw, c), 91, dtype, 1.7) + 0.4
    ones, zeros = torch.ones(c, device=_dev()), torch.zeros(c, device=_dev())
    y = ops.group_norm_nhwc(x, ones, zeros, 1e-6, 32)
    yg = y.view(n, hw, 32, c // 32)
    assert yg.mean(dim=(1, 3)).abs().max().item() < 1e-4
    assert (yg.var(dim=(1, 3), unbiased=False) - 1.0).abs().max().item() < 1e-3
    y2 = ops.group_norm_nhwc(3.0 * x + 5.0, ones, zeros, 1e-6, 32)
    assert (y2 - y).abs().max().item() < 2e-4
    add = _mk((n, c), 92, dtype)
    ya = ops.group_norm_nhwc(x, ones, zeros, 1e-6, 32, add_nc=add)
    yb = ops.group_norm_nhwc(x + add[:, None, :], ones, zeros, 1e-6, 32)
    assert (ya - yb).abs().max().item() < 2e-4


def test_group_norm_persistent_grid_and_counter_reuse():
    """The one-launch GroupNorm (vf_norm.cu: gn_resident_kernel / gn_fused_kernel) hands (sample, slab) items to a
    persistent grid through a ticket counter and synchronises the slabs of a sample through per-sample arrival counters
    that the kernel itself re-arms.  Here: more work items than resident CTAs (96 samples), a ragged row count (short last
    slab), a batch of one, and 45 back-to-back launches of alternating shapes -- a counter left dirty by one launch would
    deadlock or corrupt the next -- every result against torch, and repeated launches bit-identical."""
    from vface_b200 import ops
    dtype = torch.bfloat16
    shapes = [(96, 1000, 640), (1, 4096, 320), (7, 250, 1280), (96, 64, 1280), (33, 1024, 320)]
    cases = []
    for i, (n, hw, c) in enumerate(shapes):
        x = _mk((n, hw, c), 40 + i, dtype, 1.3) - 0.2
        w, b = _mk((c,), 50 + i, dtype), _mk((c,), 60 + i, dtype)
        add = _mk((n, c), 70 + i, dtype) if i % 2 == 0 else None
        xf = x.float() + (add.float()[:, None, :] if add is not None else 0.0)
        want = F.silu(F.group_norm(xf.permute(0, 2, 1), 32, w.float(), b.float(), 1e-5).permute(0, 2, 1))
        cases.append((x, w, b, add, want))
    first = {}
    for rep in range(9):
        for i, (x, w, b, add, want) in enumerate(cases):
            got = ops.group_norm_nhwc(x, w, b, 1e-5, 32, silu=True, add_nc=add)
            if rep == 0:
                first[i] = got
                err = (got.float() - want).abs().max().item()
                assert err < BF16_TOL * max(1.0, want.abs().max().item() / 2.0), (shapes[i], err)
            else:
                assert torch.equal(got, first[i]), (shapes[i], rep)
    torch.cuda.synchronize()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, BF16_TOL)])
@pytest.mark.parametrize("n,m,c", [(3, 4096, 320), (2, 1024, 640), (3, 253, 1280), (1, 7, 320), (2, 33, 64)])
def test_layer_norm_forms(dtype, tol, n, m, c):
    """Pure LayerNorm (packed multi-row kernel; 7 / 33 / 253 rows exercise the ragged tail of the rows-per-warp
    blocking), x + y, x + per-sample row bias, and both; the returned residual is the stored (rounded) sum."""
    from vface_b200 import ops
    x, y = _mk((n, m, c), 11 + m, dtype), _mk((n, m, c), 12 + m, dtype)
    w, b = _mk((c,), 13, dtype), _mk((c,), 14, dtype)
    row = _mk((n, c), 15, dtype)
    ln = lambda t: F.layer_norm(t, (c,), w.float(), b.float(), 1e-5)
    bound = lambda want: tol * max(1.0, want.abs().max().item() / 2.0)

    got = ops.add_layer_norm(x, w, b, 1e-5)
    want = ln(x.float())
    assert (got.float() - want).abs().max().item() < bound(want)

    for yy, rr in ((y, None), (None, row), (y, row)):
        res, got = ops.add_layer_norm(x, w, b, 1e-5, y=yy, row_bias=rr)
        s = x.float() + (yy.float() if yy is not None else 0.0) + (rr.float()[:, None, :] if rr is not None else 0.0)
        assert (res.float() - s).abs().max().item() < bound(s)
        # the kernel normalises what it stored, so both consumers see the same residual stream
        want = ln(res.float())
        assert (got.float() - want).abs().max().item() < bound(want)


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, BF16_TOL)])
def test_geglu_and_add_bias(dtype, tol):
    from vface_b200 import ops
    h = _mk((3, 1000, 2 * 1280), 21, dtype)
    got = ops.geglu(h)
    a, gate = h.float().chunk(2, dim=-1)
    want = a * F.gelu(gate)                                   # exact (erf) GELU, attention.py:43-45
    assert (got.float() - want).abs().max().item() < tol * max(1.0, want.abs().max().item() / 2.0)

    x, y = _mk((3, 777, 320), 22, dtype), _mk((3, 777, 320), 23, dtype)
    vec, rows = _mk((320,), 24, dtype), _mk((3, 320), 25, dtype)
    for bias, ref in ((vec, vec.float()), (rows, rows.float()[:, None, :]), (None, 0.0)):
        got = ops.add_bias(x, y, bias)
        want = x.float() + y.float() + ref
        assert (got.float() - want).abs().max().item() < tol * max(1.0, want.abs().max().item() / 2.0)
    inplace = x.clone()
    ops.add_bias(inplace, None, vec, out=inplace)
    assert (inplace.float() - (x.float() + vec.float())).abs().max().item() < tol * 4


@pytest.mark.parametrize("knob", ["VF_GN_PIPE", "VF_GN_WS"])
def test_group_norm_two_slab_pipeline_env_knob(knob):
    """VF_GN_PIPE=1 / VF_GN_WS=1 (read once per process): the resident GroupNorm with two half-size slabs per CTA in flight
    (vf_norm.cu: gn_resident2_kernel) and its warp-specialised form (gn_ws_kernel: producer / statistics / apply warps).  Same op, same bound: fp32 F.group_norm (+ SiLU) on the bf16 inputs, including the
    never-materialised concatenation, the additive vector and a batch small enough to fall back to the one-slab kernel;
    run twice to check that the counters were re-armed and that the result is bit-reproducible."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r'''
import torch, torch.nn.functional as F
from vface_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(5)
for n, hw, c1, c2, add in ((96, 4096, 320, 0, True), (96, 1024, 640, 320, False), (24, 256, 1280, 0, True), (3, 4096, 320, 0, False)):
    c = c1 + c2
    x1 = (torch.randn(n, hw, c1, generator=g) * 1.5 + 0.3).bfloat16().to(dev)
    x2 = torch.randn(n, hw, c2, generator=g).bfloat16().to(dev) if c2 else None
    w, b = torch.randn(c, generator=g).bfloat16().to(dev), torch.randn(c, generator=g).bfloat16().to(dev)
    a = torch.randn(n, c, generator=g).bfloat16().to(dev) if add else None
    got = ops.group_norm_nhwc(x1, w, b, 1e-5, 32, silu=True, add_nc=a, x2=x2)
    again = ops.group_norm_nhwc(x1, w, b, 1e-5, 32, silu=True, add_nc=a, x2=x2)
    assert torch.equal(got, again)
    xf = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], -1)
    if a is not None:
        xf = xf + a.float()[:, None, :]
    want = F.silu(F.group_norm(xf.permute(0, 2, 1), 32, w.float(), b.float(), 1e-5).permute(0, 2, 1))
    err = (got.float() - want).abs().max().item()
    assert err < 2e-2 * max(1.0, want.abs().max().item() / 2.0), (n, hw, c1, c2, err)
print("GN_PIPE_OK")
'''
    env = dict(os.environ, **{knob: "1"})
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=300, cwd=root)
    assert res.returncode == 0 and "GN_PIPE_OK" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
