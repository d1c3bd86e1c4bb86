"""The tcgen05 projection kernel (csrc/vf_gemm3.cu, ops.linear_proj) against an fp64 restatement of what the reference
executes around each projection of the 64x64 transformer level: LayerNorm -> Linear (ldm/modules/attention.py:239 with
:172-174), Linear + bias + per-sample row + residual (attention.py:176, :239-241), the 1x1 convolutions proj_in / proj_out
(+ x_in) (attention.py:261-288).  Tolerance: 2e-2 absolute on O(1) outputs (the bf16 bound of the other kernels); the
error actually seen is one bf16 rounding of the output (<= 2^-9 relative)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

BF16_TOL = 2e-2


def _dev():
    return torch.device("cuda:0")


def _mk(shape, seed, scale=1.0, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale + shift).to(torch.bfloat16).to(_dev())


def _ref(x, w, bias, residual, ln, row_bias):
    xd = x.double()
    if ln is not None:
        mu = xd.mean(-1, keepdim=True)
        var = xd.var(-1, unbiased=False, keepdim=True)
        xd = (xd - mu) / torch.sqrt(var + ln.eps) * ln.weight.double() + ln.bias.double()
    y = xd @ w.double().t()
    if bias is not None:
        y = y + bias.double()
    if row_bias is not None:
        y = y + row_bias.double()[:, None, :]
    if residual is not None:
        y = y + residual.double()
    return y


def _ln(k, seed):
    ln = torch.nn.LayerNorm(k).to(_dev()).to(torch.bfloat16)
    with torch.no_grad():
        ln.weight.copy_(_mk((k,), seed, 0.3, 1.0))
        ln.bias.copy_(_mk((k,), seed + 1, 0.2))
    return ln


@pytest.mark.parametrize("batch,tokens,k,n", [(3, 4096, 320, 320), (2, 4096, 320, 960), (1, 1024, 320, 160),
                                             (2, 256, 64, 320), (5, 384, 192, 480)])
@pytest.mark.parametrize("with_ln,with_bias,with_res,with_row", [
    (False, False, False, False), (True, False, False, False), (False, True, True, True), (True, True, False, False),
    (False, True, False, False), (False, False, True, False)])
def test_linear_proj_vs_fp64(batch, tokens, k, n, with_ln, with_bias, with_res, with_row):
    from vface_b200 import ops
    x = _mk((batch, tokens, k), 1 + k + n, 1.5, 0.4)
    w = _mk((n, k), 2, k ** -0.5)
    bias = _mk((n,), 3, 0.5) if with_bias else None
    res = _mk((batch, tokens, n), 4, 2.0) if with_res else None
    row = _mk((batch, n), 5, 0.7) if with_row else None
    ln = _ln(k, 6) if with_ln else None
    assert ops.linear_proj_supported(x, w)
    got = ops.linear_proj(x, w, bias, res, ln=ln, row_bias=row)
    want = _ref(x, w, bias, res, ln, row)
    assert got.shape == want.shape and got.dtype == torch.bfloat16
    err = (got.double() - want).abs().max().item()
    assert err < BF16_TOL * max(1.0, want.abs().max().item() / 2.0), err
    # and no worse than rounding the exact result to bf16 would explain, three times over
    rel = ((got.double() - want).norm() / want.norm()).item()
    assert rel < 3 * 2.0 ** -9, rel


def test_linear_proj_ragged_rows_and_untouched_tail():
    """rows not a multiple of the 128-row tile: the last tile's loads are zero-filled, its stores clipped (the rows behind
    the tensor keep their guard pattern)."""
    from vface_b200 import ops, _lib
    rows, k, n = 1000, 320, 320
    x = _mk((rows, k), 11)
    w = _mk((n, k), 12, k ** -0.5)
    r = _mk((rows, n), 13)
    ln = _ln(k, 14)
    got_ln = ops.linear_proj(x, w, None, None, ln=ln)
    want_ln = _ref(x, w, None, None, ln, None)
    assert (got_ln.double() - want_ln).abs().max().item() < BF16_TOL * max(1.0, want_ln.abs().max().item() / 2.0)
    got = ops.linear_proj(x, w, None, r)
    want = _ref(x, w, None, r, None, None)
    assert (got.double() - want).abs().max().item() < BF16_TOL * max(1.0, want.abs().max().item() / 2.0)
    # guard rows: call the C-ABI on a view of a larger buffer
    buf = torch.full((rows + 64, n), 7.0, dtype=torch.bfloat16, device=_dev())
    lib = _lib.load()
    rc = lib.vf_linear_proj(x.data_ptr(), w.data_ptr(), None, 0, r.data_ptr(), None, 0.0, None, 0, None,
                            buf.data_ptr(), rows, k, n, k, n, n, _lib.VF_BF16, torch.cuda.current_stream().cuda_stream)
    assert rc == 0, lib.vf_last_error()
    torch.cuda.synchronize()
    assert torch.equal(buf[:rows], got) and bool((buf[rows:] == 7.0).all())


def test_linear_proj_refuses_ln_with_residual():
    from vface_b200 import ops, _lib
    x, w, r = _mk((256, 320), 1), _mk((320, 320), 2), _mk((256, 320), 3)
    with pytest.raises((RuntimeError, ValueError)):
        ops.linear_proj(x, w, None, r, ln=_ln(320, 4))
    # and the C-ABI itself is loud about shapes and operands it does not take
    lib = _lib.load()
    out = torch.empty(256, 320, dtype=torch.bfloat16, device=_dev())
    b32 = torch.zeros(2, 320, device=_dev())
    st = torch.cuda.current_stream().cuda_stream
    call = lambda **kw: lib.vf_linear_proj(x.data_ptr(), w.data_ptr(), kw.get("bias"), kw.get("rpb", 0), kw.get("res"), kw.get("cs"), 1e-5,
                                           None, 0, None, out.data_ptr(), kw.get("rows", 256), kw.get("k", 320), kw.get("n", 320),
                                           320, 320, 320, kw.get("dtype", _lib.VF_BF16), st)
    assert call() == 0
    assert call(k=300) != 0 and b"bad shape" in lib.vf_last_error()
    assert call(n=200) != 0
    assert call(dtype=_lib.VF_F32) != 0
    assert call(bias=b32.data_ptr(), rpb=100) != 0            # not a multiple of the 128-row tile
    assert call(bias=b32.data_ptr(), rpb=384) != 0            # does not divide the row count
    assert call(res=out.data_ptr()) != 0                      # out aliases residual
    torch.cuda.synchronize()


def test_linear_proj_residual_is_added_exactly():
    """The residual enters the fp32 accumulator through MMAs against an identity: with a zero weight the output is the
    residual bit for bit (every column block of the 160-wide slice, including the overlapping third box)."""
    from vface_b200 import ops
    x = _mk((3, 256, 320), 41)
    r = _mk((3, 256, 960), 42, 3.0)
    w = torch.zeros(960, 320, dtype=torch.bfloat16, device=_dev())
    assert torch.equal(ops.linear_proj(x, w, None, r), r)


@pytest.mark.parametrize("rows", [4096, 1000])
def test_linear_proj_row_statistics_hand_over(rows):
    """proj_in's epilogue emits {sum, sum of squares} of its bf16 output rows per 160-column slice; the next projection's
    LayerNorm takes mean / rstd from them (SpatialTransformer -> BasicTransformerBlock.norm1 -> to_q/k/v,
    attention.py:279,239,172-174).  The emitted numbers are exact sums of the stored values (fp32 accumulation), and the
    handed-over form matches the fp64 reference as well as the in-kernel form does."""
    from vface_b200 import ops
    g = _mk((rows, 320), 51, 1.2, 0.3)
    w_in, b_in = _mk((320, 320), 52, 320 ** -0.5), _mk((320,), 53, 0.4)
    t, st = ops.linear_proj(g, w_in, b_in, emit_stats=True)
    assert st.shape == (rows, 2, 2) and torch.equal(t, ops.linear_proj(g, w_in, b_in))
    td = t.double()
    want_sum = torch.stack([td[:, :160].sum(1), td[:, 160:].sum(1)], 1)
    want_sq = torch.stack([(td[:, :160] ** 2).sum(1), (td[:, 160:] ** 2).sum(1)], 1)
    assert (st[..., 0].double() - want_sum).abs().max().item() < 1e-3
    assert ((st[..., 1].double() - want_sq).abs() / want_sq).max().item() < 1e-5
    ln = _ln(320, 54)
    wq = _mk((960, 320), 55, 320 ** -0.5)
    got = ops.linear_proj(t, wq, ln=ln, ln_stats=st)
    ref = _ref(t, wq, None, None, ln, None)
    rel = ((got.double() - ref).norm() / ref.norm()).item()
    assert rel < 3 * 2.0 ** -9, rel
    own = ops.linear_proj(t, wq, ln=ln)
    assert (got.float() - own.float()).abs().max().item() <= 2.0 ** -7 * max(1.0, ref.abs().max().item())


def test_linear_proj_large_mean_rows():
    """The folded LayerNorm subtracts mean * colsum AFTER the fp32 accumulation: rows whose mean is 20x their spread
    still come out at bf16 accuracy."""
    from vface_b200 import ops
    x = _mk((2, 512, 320), 21, 0.5, 10.0)
    w = _mk((320, 320), 22, 320 ** -0.5)
    ln = _ln(320, 23)
    got = ops.linear_proj(x, w, None, None, ln=ln)
    want = _ref(x, w, None, None, ln, None)
    rel = ((got.double() - want).norm() / want.norm()).item()
    assert rel < 3 * 2.0 ** -9, rel


def test_linear_proj_full_size_matches_library_path():
    """At the size of a 32-frame step (96 x 4096 rows): every CTA walks ~40 row blocks through both accumulator sets and
    all ring phases; compared with LayerNorm kernel + library GEMM (+ residual) of the first version of this path."""
    import torch.nn.functional as F
    from vface_b200 import ops
    b, t, k = 96, 4096, 320
    g = torch.Generator(device="cuda").manual_seed(5)
    x = (torch.randn(b, t, k, device="cuda", generator=g) * 1.3 + 0.2).bfloat16()
    ln = _ln(k, 31)
    wq = (torch.randn(960, k, device="cuda", generator=g) * k ** -0.5).bfloat16()
    got = ops.linear_proj(x, wq, None, None, ln=ln)
    want = F.linear(ops.add_layer_norm(x, ln.weight, ln.bias, ln.eps), wq)
    d = (got.float() - want.float())
    assert d.abs().max().item() < BF16_TOL * max(1.0, want.float().abs().max().item() / 2.0)
    assert (d.norm() / want.float().norm()).item() < 4 * 2.0 ** -9
    wo = (torch.randn(320, k, device="cuda", generator=g) * k ** -0.5).bfloat16()
    bo = (torch.randn(320, device="cuda", generator=g) * 0.3).bfloat16()
    row = (torch.randn(b, 320, device="cuda", generator=g) * 0.3).bfloat16()
    a = torch.randn(b, t, k, device="cuda", generator=g).bfloat16()
    got2 = ops.linear_proj(a, wo, bo, x, row_bias=row)
    want2 = ops.linear_residual(a, wo, (bo.float()[None] + row.float()).bfloat16().contiguous(), x)
    d2 = got2.float() - want2.float()
    assert (d2.norm() / want2.float().norm()).item() < 4 * 2.0 ** -9
    assert d2.abs().max().item() < BF16_TOL * max(1.0, want2.float().abs().max().item() / 2.0)


def test_linear_proj_vs_reference_golden():
    """The four uses of the kernel against outputs of the reference's own modules (tests/golden/block_projection.npz:
    BasicTransformerBlock.norm1 + attn1.to_q/k/v, attn1.to_out + x + single-token attn2, SpatialTransformer.proj_in and
    proj_out + x_in, computed by the unmodified reference in fp32).  bf16 tolerance 2e-2 on O(1..4) values."""
    import os
    import numpy as np
    from oracle.make_golden import block_projection_inputs
    from vface_b200 import ops
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "block_projection.npz"))
    d = {k: v.to(_dev()).bfloat16() for k, v in block_projection_inputs().items()}      # all values are bf16-representable
    ln = torch.nn.LayerNorm(320).to(_dev()).bfloat16()
    with torch.no_grad():
        ln.weight.copy_(d["n1_w"])
        ln.bias.copy_(d["n1_b"])

    def close(got, name):
        want = torch.from_numpy(g[name]).to(_dev())
        err = (got.float() - want).abs().max().item()
        assert err < BF16_TOL * max(1.0, want.abs().max().item() / 2.0), (name, err)
        assert ((got.float() - want).norm() / want.norm()).item() < 3 * 2.0 ** -9, name

    wqkv = torch.cat([d["wq"], d["wk"], d["wv"]], 0).contiguous()
    close(ops.linear_proj(d["x"], wqkv, ln=ln), "qkv")
    t, st = ops.linear_proj(d["g_in"], d["w_in"], d["b_in"], emit_stats=True)
    close(t, "proj_in")
    row = torch.nn.functional.linear(torch.nn.functional.linear(d["ctx"][:, 0], d["wv2"]), d["wo2"], d["bo2"])
    close(ops.linear_proj(d["a"], d["wo1"], d["bo1"], d["x"], row_bias=row), "x2")
    close(ops.linear_proj(d["t_out"], d["w_out"], d["b_out"], d["x"]), "proj_out")
    # and the hand-over: statistics emitted for g_in's projection drive the LayerNorm of a projection of that output
    q2 = ops.linear_proj(t, wqkv, ln=ln, ln_stats=st)
    assert (q2.float() - ops.linear_proj(t, wqkv, ln=ln).float()).abs().max().item() < BF16_TOL
