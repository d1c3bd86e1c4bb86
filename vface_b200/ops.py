"""Tensor-level entry points of the four hot-path kernels.

Every function takes CUDA torch tensors, passes raw device pointers and the current stream through
the C-ABI (include/vface_b200.h) and returns torch tensors.  PyTorch is only the allocator and stream
provider here; there is no eager/CPU fallback -- CPU tensors or a missing library raise.
"""
from __future__ import annotations

import os
import threading

from typing import Optional, Sequence, Union

import torch

from . import _lib
from ._lib import VF_BF16, VF_F32

# number of kernels launched through the C-ABI since import (bench.py reports it as gpu_launches)
launch_count = 0


def _count(n=1):
    global launch_count
    launch_count += n


# ---- live kernel timing (bench.py: `roofline` and `roofline_secondary`) -----------------------------------
_prof = None        # None, or {"filter": callable(kind, meta) -> bool, "events": [(kind, work, unit, ev0, ev1)]}


def profile_kernels(start=None, only=None):
    """profile_kernels(True) starts recording a CUDA event pair (on the launching stream) around every C-ABI launch
    whose op describes itself with _describe(); `only` = iterable of kind prefixes to restrict it.
    profile_kernels(None) stops and returns {kind: [(ms, work, unit), ...]} -- work = ALGORITHMIC bytes or flops of
    that launch, not measured traffic."""
    global _prof
    if start:
        pre = tuple(only) if only else None
        _prof = {"filter": (lambda k: True) if pre is None else (lambda k: k.startswith(pre)), "events": []}
        return None
    if _prof is None:
        return {}
    torch.cuda.synchronize()
    out = {}
    for kind, work, unit, a, b in _prof["events"]:
        out.setdefault(kind, []).append((a.elapsed_time(b), work, unit))
    _prof = None
    return out


def profile_attention(min_tokens):
    """Live timing of the attention kernel for the roofline in bench.py: profile_attention(n) starts recording around
    every vf_attn_fwd launch with n_q >= n; profile_attention(None) stops and returns the per-launch durations in ms."""
    if min_tokens is not None:
        profile_kernels(True, only=[f"attn n>={int(min_tokens)}"])
        _prof["attn_min"] = int(min_tokens)
        return None
    res = profile_kernels(None)
    return [ms for v in res.values() for ms, _, _ in v]


def _describe(kind: str, work: float, unit: str = "B"):
    """Called by an op right before its C-ABI launch: what the launch is and its algorithmic work."""
    _op.desc = (kind, float(work), unit)


def _code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return VF_F32
    if t.dtype == torch.bfloat16:
        return VF_BF16
    hint = ""
    if t.dtype == torch.float16:
        hint = (" (float16 usually means a torch op ran under torch.autocast: the mirrored modules switch autocast off "
                "inside their forward; wrap foreign callers in torch.autocast('cuda', enabled=False))")
    raise TypeError(f"vface_b200 kernels take float32 or bfloat16 tensors, got {t.dtype}{hint}")


_op = threading.local()     # device index of the operands of the op being issued (set by _need_cuda)


def _need_cuda(*ts):
    """Every operand on ONE CUDA device; that device is what the C-ABI call will run on (see _Lib)."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("vface_b200 ops need CUDA tensors: there is no CPU fallback")
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"vface_b200 ops need all operands on one device, got {dev} and {t.device}")
    _op.index = dev.index if dev is not None else None


class _Lib:
    """The C-ABI entry points, called with the operands' device current.

    The C side launches on the CURRENT device (and validates it: one process drives one GPU, include/vface_b200.h),
    while the stream handed over belongs to the operands' device; a tensor on cuda:1 while cuda:0 is current would
    otherwise launch on the wrong device with a foreign stream."""

    def __getattr__(self, name):
        fn = getattr(_lib.load(), name)

        def call(*args):
            idx = getattr(_op, "index", None)
            if idx is not None and idx != torch.cuda.current_device():
                with torch.cuda.device(idx):
                    return call(*args)
            desc = getattr(_op, "desc", None)
            if desc is not None:
                _op.desc = None
                if _prof is not None and _prof["filter"](desc[0]):
                    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ev0.record()
                    rc = fn(*args)
                    ev1.record()
                    _prof["events"].append((desc[0], desc[1], desc[2], ev0, ev1))
                    return rc
            return fn(*args)

        setattr(self, name, call)        # resolved once per symbol
        return call


_LIB = _Lib()


def _stream(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream


def _rows3(t: torch.Tensor, name: str):
    """(batch, n, c) tensor with unit channel stride and uniform row stride -> (ld, batch, n, c)."""
    if t.dim() != 3:
        raise ValueError(f"{name}: expected (batch, n, c), got {tuple(t.shape)}")
    b, n, c = t.shape
    if t.stride(2) != 1 or (b > 1 and t.stride(0) != n * t.stride(1)):
        raise ValueError(f"{name}: need unit channel stride and batch stride == n * row stride, got strides {t.stride()}")
    return t.stride(1), b, n, c


def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              k2: Optional[torch.Tensor] = None, v2: Optional[torch.Tensor] = None,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """softmax(q k^T * scale) v per head on the (batch, n, heads*d) layout
    (ldm/models/pnp_utils.py:270-286).  k2/v2: optional concatenated key/value segment."""
    _need_cuda(q, k, v, k2, v2, out)
    ld_q, b, n_q, c = _rows3(q, "q")
    ld_k, bk, n_kv, ck = _rows3(k, "k")
    ld_v, bv, n_v, cv = _rows3(v, "v")
    if (bk, ck) != (b, c) or (bv, n_v, cv) != (b, n_kv, c) or c % heads:
        raise ValueError(f"attention: inconsistent shapes q{tuple(q.shape)} k{tuple(k.shape)} v{tuple(v.shape)} heads={heads}")
    if not (q.dtype == k.dtype == v.dtype):
        raise TypeError("attention: q, k, v must share a dtype")
    d = c // heads
    if scale is None:
        scale = d ** -0.5
    if out is None:
        out = torch.empty((b, n_q, c), dtype=q.dtype, device=q.device)
    ld_o, bo, no, co = _rows3(out, "out")
    if (bo, no, co) != (b, n_q, c) or out.dtype != q.dtype:
        raise ValueError("attention: bad out tensor")
    p_k2 = p_v2 = None
    n_kv2 = ld_k2 = ld_v2 = 0
    if k2 is not None:
        ld_k2, b2, n_kv2, c2 = _rows3(k2, "k2")
        ld_v2, b3, n3, c3 = _rows3(v2, "v2")
        if (b2, c2) != (b, c) or (b3, n3, c3) != (b, n_kv2, c) or k2.dtype != q.dtype or v2.dtype != q.dtype:
            raise ValueError("attention: bad k2/v2")
        p_k2, p_v2 = k2.data_ptr(), v2.data_ptr()
    lib = _LIB
    tag = ""
    if _prof is not None and n_q >= _prof.get("attn_min", 1 << 60):
        tag = f"attn n>={_prof['attn_min']}"
    _describe(tag or f"attn n={n_q} h={heads} d={d}", 4.0 * n_q * (n_kv + n_kv2) * d * heads * b, "flop")
    rc = lib.vf_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), b, heads, n_q, n_kv, d,
                         ld_q, ld_k, ld_v, ld_o, float(scale), p_k2, p_v2, n_kv2, ld_k2, ld_v2,
                         _code(q), _stream(q))
    _lib.check(rc, "vf_attn_fwd")
    _count()
    return out


def fsai_blend(donor: torch.Tensor, dst: torch.Tensor, split_ratio: float = 0.8,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """combine_fft_high_low(q1=donor, q2=dst, split_ratio) (scripts/face_swap_utils.py:425-464).
    out=None returns a new tensor; out=dst reproduces the reference's in-place slice assignment."""
    _need_cuda(donor, dst, out)
    ld_d, b, n, d = _rows3(donor, "donor")
    ld_s, b2, n2, d2 = _rows3(dst, "dst")
    if (b, n, d) != (b2, n2, d2) or donor.dtype != dst.dtype:
        raise ValueError("fsai_blend: donor/dst mismatch")
    if out is None:
        out = torch.empty((b, n, d), dtype=dst.dtype, device=dst.device)
    ld_o, b3, n3, d3 = _rows3(out, "out")
    if (b3, n3, d3) != (b, n, d) or out.dtype != dst.dtype:
        raise ValueError("fsai_blend: bad out tensor")
    split = int(d * split_ratio)
    lib = _LIB
    _describe(f"fsai_blend d={d}", 3.0 * b * n * d * dst.element_size())
    rc = lib.vf_fsai_blend(donor.data_ptr(), dst.data_ptr(), out.data_ptr(), b * n, d, split,
                           ld_d, ld_s, ld_o, _code(dst), _stream(dst))
    _lib.check(rc, "vf_fsai_blend")
    _count()
    return out


def fsai_blend2(donor: torch.Tensor, dst_a: torch.Tensor, dst_b: torch.Tensor, split_ratio: float = 0.8,
                out_a: Optional[torch.Tensor] = None, out_b: Optional[torch.Tensor] = None):
    """Both branches of one tensor in one launch (ldm/models/pnp_utils.py:195+198 or :196+199):
    out_a = FSAI(donor, dst_a), out_b = FSAI(donor, dst_b); outputs default to in place."""
    _need_cuda(donor, dst_a, dst_b, out_a, out_b)
    out_a = dst_a if out_a is None else out_a
    out_b = dst_b if out_b is None else out_b
    ld_d, b, n, d = _rows3(donor, "donor")
    lds = []
    for name, t in (("dst_a", dst_a), ("out_a", out_a), ("dst_b", dst_b), ("out_b", out_b)):
        ld, bb, nn, dd = _rows3(t, name)
        if (bb, nn, dd) != (b, n, d) or t.dtype != donor.dtype:
            raise ValueError(f"fsai_blend2: {name} mismatch")
        lds.append(ld)
    split = int(d * split_ratio)
    lib = _LIB
    _describe(f"fsai_blend2 d={d}", 5.0 * b * n * d * donor.element_size())
    rc = lib.vf_fsai_blend2(donor.data_ptr(), dst_a.data_ptr(), out_a.data_ptr(), dst_b.data_ptr(), out_b.data_ptr(),
                            b * n, d, split, ld_d, lds[0], lds[1], lds[2], lds[3], _code(donor), _stream(donor))
    _lib.check(rc, "vf_fsai_blend2")
    _count()
    return out_a, out_b


def _as_flow(flow: Union[torch.Tensor, Sequence[torch.Tensor]], device) -> torch.Tensor:
    if isinstance(flow, (list, tuple)):
        flow = torch.cat([f.reshape(1, 2, f.shape[-2], f.shape[-1]) for f in flow], dim=0)
    return flow.to(device=device, dtype=torch.float32).contiguous()


def flow_warp_blend(x: torch.Tensor, flow, alpha: float, h: int, w: int,
                    prev_halo: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None,
                    return_taps: bool = False):
    """align_by_flow (scripts/temporal_flow.py:222-237) on the token layout (frames, h*w, c).
    flow: (n_flow, 2, h, w) fp32 tensor or the reference's list of (1, 2, h, w) tensors."""
    _need_cuda(x, prev_halo, out)
    ld_x, frames, n, c = _rows3(x, "x")
    if n != h * w:
        raise ValueError(f"flow_warp_blend: n={n} != h*w={h * w}")
    n_flow = frames if prev_halo is not None else frames - 1
    fl = None
    if n_flow > 0:
        fl = _as_flow(flow, x.device)
        if tuple(fl.shape) != (n_flow, 2, h, w):
            raise ValueError(f"flow_warp_blend: flow shape {tuple(fl.shape)} != {(n_flow, 2, h, w)} "
                             "(flow must be at feature resolution, SURVEY.md F5)")
    if out is None:
        out = torch.empty((frames, n, c), dtype=x.dtype, device=x.device)
    ld_o, fo, no, co = _rows3(out, "out")
    if (fo, no, co) != (frames, n, c) or out.dtype != x.dtype:
        raise ValueError("flow_warp_blend: bad out tensor")
    ld_h = 0
    p_h = None
    if prev_halo is not None:
        if prev_halo.dim() != 2 or tuple(prev_halo.shape) != (n, c) or prev_halo.stride(1) != 1 or prev_halo.dtype != x.dtype:
            raise ValueError("flow_warp_blend: prev_halo must be (h*w, c) with unit channel stride")
        ld_h = prev_halo.stride(0)
        p_h = prev_halo.data_ptr()
    taps = None
    if return_taps:
        taps = torch.empty((max(n_flow, 1), n, 2), dtype=torch.int32, device=x.device)
    lib = _LIB
    _describe(f"flow_warp c={c} hw={h}x{w}", (3.0 * n_flow + 2.0 * (frames - n_flow)) * n * c * x.element_size() + 8.0 * n * n_flow)
    rc = lib.vf_flow_warp_blend(x.data_ptr(), p_h, fl.data_ptr() if fl is not None else None, out.data_ptr(),
                                frames, h, w, c, ld_x, ld_h, ld_o, float(alpha), _code(x),
                                taps.data_ptr() if taps is not None else None, _stream(x))
    _lib.check(rc, "vf_flow_warp_blend")
    _count()
    return (out, taps) if return_taps else out


def ddim_cfg_step(x: torch.Tensor, e_uncond: torch.Tensor, e_cond: torch.Tensor,
                  a_t: float, a_prev: float, sigma_t: float, sqrt_one_minus_at: float, cfg_scale: float,
                  noise: Optional[torch.Tensor] = None):
    """CFG + DDIM update (ldm/models/diffusion/ddim_w_inv.py:666, :686, :696-700) -> (x_prev, pred_x0)."""
    _need_cuda(x, e_uncond, e_cond, noise)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("ddim_cfg_step: x must be contiguous float32")
    if not (e_uncond.is_contiguous() and e_cond.is_contiguous()) or e_uncond.dtype != e_cond.dtype:
        raise ValueError("ddim_cfg_step: e_uncond/e_cond must be contiguous and share a dtype")
    if e_uncond.numel() != x.numel() or e_cond.numel() != x.numel():
        raise ValueError("ddim_cfg_step: size mismatch")
    if noise is not None and (noise.dtype != torch.float32 or not noise.is_contiguous() or noise.numel() != x.numel()):
        raise ValueError("ddim_cfg_step: noise must be contiguous float32 of x's size")
    x_prev = torch.empty_like(x)
    pred_x0 = torch.empty_like(x)
    lib = _LIB
    _describe("ddim_cfg_step", x.numel() * (3.0 * 4 + 2.0 * e_cond.element_size() + (4.0 if noise is not None else 0.0)))
    rc = lib.vf_ddim_cfg_step(x.data_ptr(), e_uncond.data_ptr(), e_cond.data_ptr(), x_prev.data_ptr(), pred_x0.data_ptr(),
                              float(a_t), float(a_prev), float(sigma_t), float(sqrt_one_minus_at), float(cfg_scale),
                              noise.data_ptr() if noise is not None else None, x.numel(), _code(e_cond), _stream(x))
    _lib.check(rc, "vf_ddim_cfg_step")
    _count()
    return x_prev, pred_x0


def ddim_invert_step(x: torch.Tensor, e_cond: torch.Tensor, a_cur: float, a_next: float,
                     e_uncond: Optional[torch.Tensor] = None, cfg_scale: float = 1.0) -> torch.Tensor:
    """Forward-DDIM update of ddim_invert (ldm/models/diffusion/ddim_w_inv.py:426-449)."""
    _need_cuda(x, e_cond, e_uncond)
    if x.dtype != torch.float32 or not x.is_contiguous() or not e_cond.is_contiguous() or e_cond.numel() != x.numel():
        raise ValueError("ddim_invert_step: x must be contiguous float32 and e_cond contiguous of the same size")
    if e_uncond is not None and (not e_uncond.is_contiguous() or e_uncond.dtype != e_cond.dtype or e_uncond.numel() != x.numel()):
        raise ValueError("ddim_invert_step: bad e_uncond")
    x_next = torch.empty_like(x)
    lib = _LIB
    rc = lib.vf_ddim_invert_step(x.data_ptr(), e_uncond.data_ptr() if e_uncond is not None else None, e_cond.data_ptr(),
                                 x_next.data_ptr(), float(a_cur), float(a_next), float(cfg_scale), x.numel(),
                                 _code(e_cond), _stream(x))
    _lib.check(rc, "vf_ddim_invert_step")
    _count()
    return x_next


# ---- fused glue kernels of the UNet (vf_norm.cu) ---------------------------------------------------------
_GN_FUSED = os.environ.get("VF_GN_FUSED", "1") != "0"
def _rows_c(t: torch.Tensor, name: str):
    if not t.is_contiguous():
        raise ValueError(f"{name}: expected a contiguous (..., c) tensor, got strides {t.stride()}")
    c = t.shape[-1]
    return t.numel() // c, c


def group_norm_nhwc(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float, groups: int = 32,
                    silu: bool = False, add_nc: Optional[torch.Tensor] = None,
                    x2: Optional[torch.Tensor] = None) -> torch.Tensor:
    """GroupNorm (+SiLU) of channels-last activations x: (n, ..., c); add_nc: optional (n, c) vector
    added to every position before the statistics (ResBlock `h + emb_out`).
    x2: optional second channels-last tensor (n, ..., c2): the norm runs over the channel concatenation
    [x ; x2] read in place (the skip-connection `th.cat` of the UNet decoder, openaimodel.py:899) and
    returns the normalised (n, ..., c + c2) tensor."""
    _need_cuda(x, weight, bias, add_nc, x2)
    if not x.is_contiguous() or x.dim() < 3:
        raise ValueError("group_norm_nhwc: x must be a contiguous (n, ..., c) tensor")
    n, c1 = x.shape[0], x.shape[-1]
    hw = x.numel() // (n * c1)
    c2 = 0
    if x2 is not None:
        if not x2.is_contiguous() or x2.shape[:-1] != x.shape[:-1] or x2.dtype != x.dtype:
            raise ValueError("group_norm_nhwc: x2 must be contiguous and match x in all but the channel axis")
        c2 = x2.shape[-1]
    c = c1 + c2
    if weight.dtype != x.dtype or bias.dtype != x.dtype or weight.numel() != c or bias.numel() != c or \
            (add_nc is not None and (add_nc.dtype != x.dtype or tuple(add_nc.shape) != (n, c))):
        raise ValueError("group_norm_nhwc: parameter dtype/shape mismatch")
    lib = _LIB
    ws = torch.empty(lib.vf_group_norm_workspace_floats(n, hw, groups), dtype=torch.float32, device=x.device)
    y = torch.empty(x.shape[:-1] + (c,), dtype=x.dtype, device=x.device)
    _describe(f"group_norm c={c}" + ("+silu" if silu else ""), 3.0 * n * hw * c * x.element_size())
    rc = lib.vf_group_norm_nhwc_cat(x.data_ptr(), c1, x2.data_ptr() if x2 is not None else None, c2,
                                    add_nc.contiguous().data_ptr() if add_nc is not None else None,
                                    weight.data_ptr(), bias.data_ptr(), y.data_ptr(), ws.data_ptr(), n, hw, groups,
                                    float(eps), int(silu), _code(x), _stream(x))
    _lib.check(rc, "vf_group_norm_nhwc_cat")
    # one persistent launch (vf_norm.cu: gn_fused_kernel) unless the row is wider than 320 16-byte chunks, the batch
    # exceeds its counter table, or VF_GN_FUSED=0 keeps the statistics / apply pair
    fused = _GN_FUSED and c * x.element_size() // 16 <= 320 and n <= 4096
    _count(1 if fused else 2)
    return y


def add_layer_norm(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, eps: float = 1e-5,
                   y: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None):
    """res = x + y + row_bias ; out = LayerNorm(res).  row_bias: (c,) or (n, c) for x of shape (n, m, c).
    Returns out when neither y nor row_bias is given, else (res, out)."""
    _need_cuda(x, weight, bias, y, row_bias)
    rows, c = _rows_c(x, "x")
    if y is not None and (y.shape != x.shape or y.dtype != x.dtype or not y.is_contiguous()):
        raise ValueError("add_layer_norm: y must match x and be contiguous")
    rpb = 0
    if row_bias is not None:
        row_bias = row_bias.contiguous()
        if row_bias.dtype != x.dtype or row_bias.shape[-1] != c:
            raise ValueError("add_layer_norm: bad row_bias")
        nb = row_bias.numel() // c
        if nb > 1:
            if rows % nb:
                raise ValueError("add_layer_norm: rows not divisible by bias rows")
            rpb = rows // nb
    fused = y is not None or row_bias is not None
    res = torch.empty_like(x) if fused else None
    out = torch.empty_like(x)
    lib = _LIB
    _describe(f"layer_norm c={c}" + ("+add" if fused else ""), (2.0 + (1.0 if y is not None else 0.0) + (1.0 if fused else 0.0)) * rows * c * x.element_size())
    rc = lib.vf_add_layer_norm(x.data_ptr(), y.data_ptr() if y is not None else None,
                               row_bias.data_ptr() if row_bias is not None else None, rpb, weight.data_ptr(), bias.data_ptr(),
                               res.data_ptr() if fused else None, out.data_ptr(), rows, c, float(eps), _code(x), _stream(x))
    _lib.check(rc, "vf_add_layer_norm")
    _count()
    return (res, out) if fused else out


def geglu(h: torch.Tensor) -> torch.Tensor:
    """h (..., 2k) -> h[..., :k] * gelu(h[..., k:])  (exact erf GELU, attention.py:43-45)."""
    _need_cuda(h)
    rows, two_k = _rows_c(h, "h")
    k = two_k // 2
    out = torch.empty(h.shape[:-1] + (k,), dtype=h.dtype, device=h.device)
    lib = _LIB
    _describe(f"geglu k={k}", 3.0 * rows * k * h.element_size())
    rc = lib.vf_geglu(h.data_ptr(), out.data_ptr(), rows, k, two_k, _code(h), _stream(h))
    _lib.check(rc, "vf_geglu")
    _count()
    return out


def linear_geglu_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes/dtypes the fused tcgen05 GEMM+GEGLU kernel takes (everything else: library GEMM + geglu)."""
    two_n, k = weight.shape
    return (x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.is_cuda and k % 64 == 0
            and two_n % 256 == 0 and x.shape[-1] == k and os.environ.get("VF_FUSED_GEGLU", "1") != "0")


def linear_geglu(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor]) -> torch.Tensor:
    """GEGLU.forward (ldm/modules/attention.py:37-45) as ONE kernel: (x @ Wv^T + bv) * gelu(x @ Wg^T + bg) with
    weight = [Wv ; Wg] (2n, k); the (rows, 2n) projection never reaches HBM."""
    _need_cuda(x, weight, bias)
    if not x.is_contiguous() or not weight.is_contiguous():
        raise ValueError("linear_geglu: x and weight must be contiguous")
    two_n, k = weight.shape
    n = two_n // 2
    rows = x.numel() // k
    if bias is not None and (bias.dtype != x.dtype or bias.numel() != two_n or not bias.is_contiguous()):
        raise ValueError("linear_geglu: bad bias")
    out = torch.empty(x.shape[:-1] + (n,), dtype=x.dtype, device=x.device)
    lib = _LIB
    _describe(f"linear_geglu k={k} n={n}", 2.0 * rows * k * two_n, "flop")
    rc = lib.vf_linear_geglu(x.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None, out.data_ptr(),
                             rows, k, n, k, _code(x), _stream(x))
    _lib.check(rc, "vf_linear_geglu")
    _count()
    return out


def upsample_nearest2x(x: torch.Tensor) -> torch.Tensor:
    """F.interpolate(x, scale_factor=2, mode="nearest") for a (n, c, h, w) tensor held channels-last; returns
    (n, c, 2h, 2w) channels-last (openaimodel.py:107-117)."""
    _need_cuda(x)
    n, c, h, w = x.shape
    xt = x.permute(0, 2, 3, 1).contiguous()
    out = torch.empty((n, 2 * h, 2 * w, c), dtype=x.dtype, device=x.device)
    lib = _LIB
    _describe(f"upsample2x c={c}", 5.0 * n * h * w * c * x.element_size())
    rc = lib.vf_upsample_nearest2x_nhwc(xt.data_ptr(), out.data_ptr(), n, h, w, c, _code(x), _stream(x))
    _lib.check(rc, "vf_upsample_nearest2x_nhwc")
    _count()
    return out.permute(0, 3, 1, 2)


def conv2d_nhwc(x: torch.Tensor, conv) -> torch.Tensor:
    """conv(x) for an nn.Conv2d with bias on channels-last activations: cuDNN convolution without bias, then the
    bias through vf_add_bias in place (ATen appends a non-vectorised broadcast add that runs at 1.6 TB/s)."""
    import torch.nn.functional as F
    y = F.conv2d(x, conv.weight, None, conv.stride, conv.padding, conv.dilation, conv.groups)
    if conv.bias is None:
        return y
    yt = y.permute(0, 2, 3, 1)
    vec = 16 // y.element_size()
    if not yt.is_contiguous() or not y.is_cuda or y.shape[1] % vec:     # e.g. the 4-channel output convolution
        return y + conv.bias.view(1, -1, 1, 1)
    add_bias(yt, None, conv.bias, out=yt)
    return y


_lt_workspace = {}


def linear_residual(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], residual: torch.Tensor) -> torch.Tensor:
    """residual + x @ weight^T + bias in ONE library GEMM (cuBLASLt: beta = 1 and the bias in the fp32 epilogue), for
    `x = ff(norm3(x)) + x` (attention.py:242) and `proj_out(x) + x_in` (attention.py:287-288): the projection is never
    rounded and written out just to be re-read by an add.  Returns a new tensor; `residual` is not modified.
    bias may be (n,) or, for x of shape (batch, rows, k), one row per batch entry: (batch, n)."""
    _need_cuda(x, weight, bias, residual)
    n, k = weight.shape
    if x.shape[-1] != k or residual.shape[-1] != n or x.shape[:-1] != residual.shape[:-1]:
        raise ValueError(f"linear_residual: shapes x {tuple(x.shape)}, weight {tuple(weight.shape)}, residual {tuple(residual.shape)}")
    if not (x.is_contiguous() and weight.is_contiguous() and residual.is_contiguous()):
        raise ValueError("linear_residual: contiguous tensors")
    if weight.dtype != x.dtype or residual.dtype != x.dtype or (bias is not None and bias.dtype != x.dtype):
        raise ValueError("linear_residual: one dtype")
    out = torch.empty_like(residual)
    ws = _lt_workspace.get(x.device)
    if ws is None:
        ws = _lt_workspace[x.device] = torch.empty(32 << 20, dtype=torch.uint8, device=x.device)
    lib = _LIB
    if bias is not None and bias.dim() == 2:
        if x.dim() != 3 or bias.shape != (x.shape[0], n) or not bias.is_contiguous():
            raise ValueError("linear_residual: a per-batch bias needs x (batch, rows, k) and bias (batch, n)")
        rc = lib.vf_linear_residual_batched(x.data_ptr(), weight.data_ptr(), bias.data_ptr(), residual.data_ptr(), out.data_ptr(),
                                            x.shape[0], x.shape[1], k, n, k, n, n, ws.data_ptr(), ws.numel(), _code(x), _stream(x))
        _lib.check(rc, "vf_linear_residual_batched")
        return out
    if bias is not None and (bias.numel() != n or not bias.is_contiguous()):
        raise ValueError("linear_residual: bad bias")
    rows = x.numel() // k
    rc = lib.vf_linear_residual(x.data_ptr(), weight.data_ptr(), bias.data_ptr() if bias is not None else None,
                                residual.data_ptr(), out.data_ptr(), rows, k, n, k, n, n, ws.data_ptr(), ws.numel(),
                                _code(x), _stream(x))
    _lib.check(rc, "vf_linear_residual")          # a library GEMM: not counted as a vface_b200 kernel launch
    return out


_PROJ_KNOB = os.environ.get("VF_PROJ_TC", "1") != "0"     # 0: keep the 64x64-level projections on the library GEMM


class _ProjFold:
    """Per-weight constants of ops.linear_proj, cached until a parameter changes: the LayerNorm's gamma folded into
    the weight columns (rounded to bf16 once), the fp32 column sums of THAT rounded matrix (so that the mean term
    cancels exactly), and beta . W^T + bias in fp32."""
    _cache = {}

    @classmethod
    def get(cls, weight, bias, ln):
        ps = [weight] + ([bias] if bias is not None else []) + ([ln.weight, ln.bias] if ln is not None else [])
        key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in ps) + (bias is None, ln is None)
        hit = cls._cache.get(key)
        if hit is not None:
            return hit[:3]
        if len(cls._cache) > 512:
            cls._cache.clear()
        with torch.no_grad():
            w32 = weight.detach().float()
            b32 = bias.detach().float() if bias is not None else None
            if ln is not None:
                wg = (w32 * ln.weight.detach().float()[None, :]).to(torch.bfloat16).contiguous()
                colsum = wg.float().sum(dim=1).contiguous()
                lb = w32 @ ln.bias.detach().float()
                b32 = lb if b32 is None else b32 + lb
            else:
                wg = weight.detach().contiguous()
                colsum = None
            b32 = b32.contiguous() if b32 is not None else None
        # the entry keeps its source tensors alive: a freed parameter's address (and version 0) could otherwise be handed to
        # a new tensor and hit this entry with stale folded constants
        cls._cache[key] = (wg, colsum, b32, tuple(ps))
        return wg, colsum, b32


def linear_proj_supported(x: torch.Tensor, weight: torch.Tensor) -> bool:
    """Shapes / dtypes the tcgen05 projection kernel takes (k a multiple of 64 up to 320, n a multiple of 160, bf16)."""
    n, k = weight.shape
    return (_PROJ_KNOB and x.is_cuda and x.dtype == torch.bfloat16 and weight.dtype == torch.bfloat16 and x.shape[-1] == k
            and k % 64 == 0 and k <= 320 and n % 160 == 0 and x.numel() // k >= 128)


def linear_proj(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None, ln=None, row_bias: Optional[torch.Tensor] = None,
                ln_stats: Optional[torch.Tensor] = None, emit_stats: bool = False):
    """LN(x) @ weight^T + bias + row_bias[sample] + residual as ONE tcgen05 kernel (csrc/vf_gemm3.cu).

    ln: a torch.nn.LayerNorm over the last dimension of x (or None); x is then the RAW input -- mean / rstd are computed
    inside the kernel, the normalised tensor never exists.  bias: (n,), the layer's own.  row_bias: (batch, n) for x of
    shape (batch, tokens, k) -- one extra row per sample (attn2's single-token output), tokens % 128 == 0.
    emit_stats: also return (rows, n / 160, 2) fp32 partial {sum, sum of squares} of the bf16 output rows; handed to the
    NEXT projection as ln_stats, its LayerNorm needs no look at x of its own (returns (out, stats)).
    Replaces norm1 + to_q/to_k/to_v, to_out + adds, proj_in, proj_out + x_in of the 64x64 level
    (ldm/modules/attention.py:172-176, 239-241, 261-288)."""
    _need_cuda(x, weight, bias, residual, row_bias)
    n, k = weight.shape
    if x.shape[-1] != k or not x.is_contiguous() or x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise ValueError("linear_proj: x must be contiguous bf16 (..., k) and weight bf16 (n, k)")
    if not weight.is_contiguous():
        raise ValueError("linear_proj: weight must be contiguous")
    rows = x.numel() // k
    if residual is not None and (residual.shape != x.shape[:-1] + (n,) or residual.dtype != x.dtype or not residual.is_contiguous()):
        raise ValueError("linear_proj: residual must be contiguous, of the output's shape and dtype")
    if ln is not None and (tuple(ln.normalized_shape) != (k,) or ln.weight is None or ln.bias is None):
        raise ValueError("linear_proj: ln must be an affine LayerNorm over the last dimension of x")
    if ln is not None and residual is not None:
        raise ValueError("linear_proj: the LayerNorm form takes no residual")
    wg, colsum, b32 = _ProjFold.get(weight, bias, ln)
    rpb = 0
    if row_bias is not None:
        if x.dim() != 3 or row_bias.shape != (x.shape[0], n):
            raise ValueError("linear_proj: row_bias needs x (batch, tokens, k) and row_bias (batch, n)")
        if x.shape[1] % 128:
            raise ValueError("linear_proj: tokens per sample must be a multiple of 128 for a per-sample row")
        rpb = x.shape[1]
        rb = row_bias.float()
        b32 = (rb + b32[None, :]).contiguous() if b32 is not None else rb.contiguous()
    parts = 0
    if ln_stats is not None:
        if ln is None or ln_stats.dtype != torch.float32 or not ln_stats.is_contiguous() or ln_stats.dim() != 3 \
                or ln_stats.shape[0] != rows or ln_stats.shape[2] != 2 or ln_stats.device != x.device:
            raise ValueError("linear_proj: ln_stats must be the (rows, parts, 2) fp32 array emitted for x, and needs ln")
        parts = ln_stats.shape[1]
    if emit_stats and ln is not None:
        raise ValueError("linear_proj: emit_stats is not available in the LayerNorm form")
    stats = torch.empty((rows, n // 160, 2), dtype=torch.float32, device=x.device) if emit_stats else None
    out = torch.empty(x.shape[:-1] + (n,), dtype=x.dtype, device=x.device)
    lib = _LIB
    units = (k + n + (n if residual is not None else 0)) * 2.0
    _describe(f"linear_proj k={k} n={n}" + ("+ln" if ln is not None else "") + ("(stats in)" if ln_stats is not None else "")
              + ("+res" if residual is not None else "") + ("+stats out" if emit_stats else ""), units * rows)
    rc = lib.vf_linear_proj(x.data_ptr(), wg.data_ptr(), b32.data_ptr() if b32 is not None else None, rpb,
                            residual.data_ptr() if residual is not None else None,
                            colsum.data_ptr() if colsum is not None else None, float(ln.eps) if ln is not None else 0.0,
                            ln_stats.data_ptr() if ln_stats is not None else None, parts,
                            stats.data_ptr() if stats is not None else None,
                            out.data_ptr(), rows, k, n, k, n, n, _code(x), _stream(x))
    _lib.check(rc, "vf_linear_proj")
    _count()
    return (out, stats) if emit_stats else out


def conv3x3_out_f32(x_nhwc: torch.Tensor, conv) -> torch.Tensor:
    """The UNet's output convolution (openaimodel.py:835/:907) with an fp32 result: x_nhwc (n, h, w, c) bf16
    channels-last tokens -> (n, 4, h, w) fp32 contiguous.  eps leaves the network unrounded because classifier-free
    guidance multiplies the rounding of a bf16 eps by 3.6 (ddim_w_inv.py:666)."""
    _need_cuda(x_nhwc, conv.weight)
    if x_nhwc.dtype != torch.bfloat16 or not x_nhwc.is_contiguous():
        raise ValueError("conv3x3_out_f32: contiguous bf16 (n, h, w, c) input")
    if conv.kernel_size != (3, 3) or conv.stride != (1, 1) or conv.padding != (1, 1) or conv.groups != 1:
        raise ValueError("conv3x3_out_f32: 3x3, stride 1, padding 1 convolution")
    n, h, w, c = x_nhwc.shape
    wt = conv.weight.detach().contiguous()                         # OIHW (tiny: 4 x c x 9)
    bias = conv.bias.detach().contiguous() if conv.bias is not None else None
    out = torch.empty((n, wt.shape[0], h, w), dtype=torch.float32, device=x_nhwc.device)
    lib = _LIB
    _describe(f"conv3x3_out c={c}", n * h * w * (c * 2.0 + wt.shape[0] * 4.0))
    rc = lib.vf_conv3x3_out_f32(x_nhwc.data_ptr(), wt.data_ptr(), bias.data_ptr() if bias is not None else None,
                                out.data_ptr(), n, h, w, c, wt.shape[0], _code(x_nhwc), _stream(x_nhwc))
    _lib.check(rc, "vf_conv3x3_out_f32")
    _count()
    return out


def add_bias(a: torch.Tensor, b: Optional[torch.Tensor] = None, row_bias: Optional[torch.Tensor] = None,
             out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a + b + row_bias over contiguous (..., c) tensors; row_bias (c,) or (n, c).  out may be `a` (in place)."""
    _need_cuda(a, b, row_bias)
    rows, c = _rows_c(a, "a")
    if b is not None and (b.shape != a.shape or b.dtype != a.dtype or not b.is_contiguous()):
        raise ValueError("add_bias: b must match a and be contiguous")
    rpb = 0
    if row_bias is not None:
        row_bias = row_bias.contiguous()
        if row_bias.dtype != a.dtype or row_bias.shape[-1] != c:
            raise ValueError("add_bias: bad row_bias")
        nb = row_bias.numel() // c
        if nb > 1:
            if rows % nb:
                raise ValueError("add_bias: rows not divisible by bias rows")
            rpb = rows // nb
    if out is None:
        out = torch.empty_like(a)
    elif out.shape != a.shape or out.dtype != a.dtype or not out.is_contiguous():
        raise ValueError("add_bias: out must match a and be contiguous")
    lib = _LIB
    _describe(f"add_bias c={c}", (2.0 + (1.0 if b is not None else 0.0)) * rows * c * a.element_size())
    rc = lib.vf_add_bias(a.data_ptr(), b.data_ptr() if b is not None else None,
                         row_bias.data_ptr() if row_bias is not None else None, rpb, out.data_ptr(), rows, c, _code(a), _stream(a))
    _lib.check(rc, "vf_add_bias")
    _count()
    return out
