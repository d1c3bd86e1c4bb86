"""Transformer blocks of the REFace UNet, backed by the vface_b200 attention kernel.

Host-side mirror of REFace/ldm/modules/attention.py: CrossAttention :152-221,
BasicTransformerBlock :224-243, SpatialTransformer :246-288, GEGLU/FeedForward :37-64.  Class names,
constructor arguments, sub-module names and therefore state-dict keys are the reference's, so
`last.ckpt` loads unchanged; the forward passes are re-designed:

  * q, k, v come from ONE projection GEMM over a cached concatenated weight and stay in the
    (batch, n, heads*d) layout -- the attention kernel indexes heads itself, so the reference's
    three 'b n (h d) -> (b h) n d' rearrange copies and the N x N `sim` / `attn` tensors never exist;
  * self-attention runs in vf_attn_fwd (tcgen05 for bf16, fp32 kernel for fp32 parity runs);
  * cross-attention against a single context token (the VFace conditioning is (B, 1, 768)) is a
    softmax over one key == 1, so it folds to to_out(to_v(context)) broadcast over tokens
    (SURVEY.md row a11); longer contexts go through the same attention kernel.
"""
from __future__ import annotations

from inspect import isfunction

import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from .diffusionmodules.util import zero_module


def exists(val):
    return val is not None


def default(val, d):
    if exists(val):
        return val
    return d() if isfunction(d) else d


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        # bf16: projection + gate in ONE tcgen05 GEMM kernel (the (rows, 2*dim_out) projection never reaches HBM);
        # fp32 / odd shapes: library GEMM, then x * gelu(gate) in one pass (reference :43-45)
        if ops.linear_geglu_supported(x, self.proj.weight):
            return ops.linear_geglu(x.contiguous(), self.proj.weight, self.proj.bias)
        return ops.geglu(self.proj(x))


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = default(dim_out, dim)
        project_in = GEGLU(dim, inner_dim) if glu else nn.Sequential(nn.Linear(dim, inner_dim), nn.GELU())
        self.net = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def forward(self, x):
        return self.net(x)


def Normalize(in_channels):
    return torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class CrossAttention(nn.Module):
    """Parameters: to_q/to_k/to_v.weight (no bias), to_out.0.{weight,bias} (reference :152-177)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0., sep_head_att=False):
        super().__init__()
        if sep_head_att:
            raise NotImplementedError("sep_head_att is not used by the VFace configuration")
        inner_dim = dim_head * heads
        context_dim = default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.dim_head = dim_head
        self.head_splits = [6, 2]
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))
        self._wqkv = None
        self._wqkv_key = None

    # -- fused projection -------------------------------------------------------------------------
    def fused_qkv_weight(self):
        """[Wq; Wk; Wv] (3*inner, query_dim), cached until any of the three parameters changes."""
        ws = (self.to_q.weight, self.to_k.weight, self.to_v.weight)
        key = tuple((w.data_ptr(), w._version, w.dtype, w.device) for w in ws)
        if self._wqkv is None or self._wqkv_key != key:
            self._wqkv = torch.cat([w.detach() for w in ws], dim=0).contiguous()
            self._wqkv_key = key
        return self._wqkv

    def project_qkv(self, x):
        """Self-attention projections in one GEMM: returns column-slice views q, k, v of a
        (batch, n, 3*inner) buffer (row stride 3*inner; the kernels take row strides).

        When the enclosing BasicTransformerBlock has armed `_pre_ln` (its norm1), `x` is the block's RAW input and the
        LayerNorm happens inside the projection kernel (ops.linear_proj: the normalised tensor never exists) or, for
        shapes that kernel does not take, right here."""
        inner = self.heads * self.dim_head
        w = self.fused_qkv_weight()
        ln = getattr(self, "_pre_ln", None)
        if ln is not None:
            object.__setattr__(self, "_pre_ln", None)       # (object.__setattr__: a Module value must not become a sub-module)
            if ops.linear_proj_supported(x, w):
                # row statistics handed over by the kernel that produced x (SpatialTransformer's proj_in), if any
                st = getattr(x, "_vf_row_stats", None) if _STATS_HANDOFF else None
                if st is not None and (st.shape[0] != x.numel() // x.shape[-1] or st.device != x.device):
                    st = None
                qkv = ops.linear_proj(x.contiguous(), w, ln=ln, ln_stats=st)
            else:
                qkv = F.linear(ops.add_layer_norm(x.contiguous(), ln.weight, ln.bias, ln.eps), w)
        elif ops.linear_proj_supported(x, w) and x.dim() == 3:
            qkv = ops.linear_proj(x.contiguous(), w)
        else:
            qkv = F.linear(x, w)
        return qkv[..., :inner], qkv[..., inner:2 * inner], qkv[..., 2 * inner:]

    def attend(self, q, k, v):
        """softmax(q k^T * scale) v, heads indexed inside the kernel (reference :270-286 / :206-220)."""
        return ops.attention(q, k, v, self.heads, self.scale)

    def project_out(self, a):
        """to_out(a) (reference :221 / pnp_utils.py:287).  When the enclosing BasicTransformerBlock has armed
        `_fused_out` (bf16, single-token context), the projection, its bias, attn2's per-sample row and the residual add
        are ONE library GEMM and the result is the updated residual stream; `_fused_out.done` tells the block so.
        A foreign `forward` that calls `self.to_out` directly simply leaves it unarmed-and-undone."""
        ep = getattr(self, "_fused_out", None)
        if ep is not None and not ep.done and a.dtype == torch.bfloat16 and a.dim() == 3 and a.shape[1] >= 256:
            lin = self.to_out[0]
            ep.done = True
            if ops.linear_proj_supported(a, lin.weight) and a.shape[1] % 128 == 0 and lin.weight.is_contiguous():
                # 64x64 level: the tcgen05 projection kernel (bias and per-sample row in fp32, residual through the MMA)
                return ops.linear_proj(a.contiguous(), lin.weight, None if ep.prebiased else lin.bias, ep.residual, row_bias=ep.row)
            bias = ep.row if ep.prebiased else (lin.bias + ep.row).contiguous()
            return ops.linear_residual(a.contiguous(), lin.weight, bias, ep.residual)
        return self.to_out(a)

    def single_token_row(self, context):
        """Cross-attention against ONE context token: softmax over one key == 1 exactly, so every query
        gets to_out(to_v(context)) -- (b, 1, query_dim), independent of x."""
        return self.to_out(self.to_v(context))

    def forward(self, x, context=None, mask=None):
        if exists(mask):
            raise NotImplementedError("attention masks are not used on the VFace hot path")
        if context is None:
            q, k, v = self.project_qkv(x)
            return self.project_out(self.attend(q, k, v))
        if context.shape[-1] == 768 * 2:
            raise NotImplementedError("split clip/landmark contexts (1536-wide) are not used by the VFace configuration")
        if context.shape[1] == 1:
            return self.single_token_row(context).expand(-1, x.shape[1], -1)
        q = self.to_q(x)
        k = self.to_k(context)
        v = self.to_v(context)
        return self.to_out(self.attend(q, k, v))


import os as _os
_FUSE_TO_OUT = _os.environ.get("VF_FUSE_TO_OUT", "1") != "0"     # tuning knob: 0 keeps to_out and the LN3 add separate
_STATS_HANDOFF = _os.environ.get("VF_PROJ_STATS", "1") != "0"    # tuning knob: 0 = norm1's statistics computed inside the QKV kernel
_FUSE_LN_QKV = _os.environ.get("VF_FUSE_LN_QKV", "1") != "0"     # tuning knob: 0 keeps norm1 as its own kernel


class _FusedOut:
    """Epilogue armed by BasicTransformerBlock around its attn1 call (see CrossAttention.project_out).
    `prebiased`: `row` already contains attn1's to_out bias (BasicTransformerBlock._row_plus_out_bias)."""
    __slots__ = ("residual", "row", "done", "prebiased")

    def __init__(self, residual, row, prebiased=False):
        self.residual, self.row, self.done, self.prebiased = residual, row, False, prebiased


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, dropout=0., context_dim=None, gated_ff=True, checkpoint=True,
                 sep_head_att=False):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout, sep_head_att=sep_head_att)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        self.checkpoint = checkpoint      # gradient checkpointing is a training device; inference ignores it

    def forward(self, x, context=None):
        return self._forward(x, context)

    def _row_plus_out_bias(self, context):
        """attn2's single-token row (to_out(to_v(ctx)), CrossAttention.single_token_row) PLUS attn1's to_out bias as ONE small
        GEMM: to_v has no bias and nothing sits between the two projections, so W = W_out2 W_v (formed in fp32, cached
        until a parameter changes) and b = b_out2 + b_out1.  Replaces two GEMMs, a bias add and a copy per block and step;
        bf16 only (one rounding of the row instead of two).  None when the modules are not the plain reference ones."""
        a1o, a2 = self.attn1.to_out[0], self.attn2
        a2o = a2.to_out[0]
        if (type(a1o) is not nn.Linear or type(a2o) is not nn.Linear or type(a2.to_v) is not nn.Linear
                or a2.to_v.bias is not None or a1o.bias is None or a2o.bias is None):
            return None
        ps = (a2.to_v.weight, a2o.weight, a2o.bias, a1o.bias)
        key = tuple((p.data_ptr(), p._version, p.dtype, p.device) for p in ps)
        hit = self.__dict__.get("_row_gemm")
        if hit is None or hit[0] != key:
            with torch.no_grad():
                w = (a2o.weight.float() @ a2.to_v.weight.float()).to(a2o.weight.dtype).contiguous()
                b = (a2o.bias.float() + a1o.bias.float()).to(a2o.bias.dtype).contiguous()
            hit = (key, w, b)
            self.__dict__["_row_gemm"] = hit
        return F.linear(context[:, 0], hit[1], hit[2])

    def _attn1_of_norm1(self, x, ln):
        """attn1(norm1(x)) (reference :239).  bf16 with this package's own attn1 forward (the class's, or the VFace hook
        closure of ldm/models/pnp_utils.py, both of which use their input only through project_qkv): norm1 is handed to
        project_qkv and folded into the QKV projection kernel; a foreign `forward` gets the normalised tensor as before."""
        a1 = self.attn1
        fwd = a1.__dict__.get("forward")
        own = fwd is None or getattr(fwd, "_vf_native", False)
        if (_FUSE_LN_QKV and own and type(a1) is CrossAttention and x.dtype == torch.bfloat16 and x.dim() == 3
                and type(self.norm1) is nn.LayerNorm and ops.linear_proj_supported(x, a1.to_q.weight)):
            object.__setattr__(a1, "_pre_ln", self.norm1)
            try:
                return a1(x)
            finally:
                object.__setattr__(a1, "_pre_ln", None)
        return a1(ln(self.norm1, x))

    def _forward(self, x, context=None):
        """x = attn1(LN1(x)) + x ; x = attn2(LN2(x), ctx) + x ; x = ff(LN3(x)) + x   (reference :239-243),
        with each residual add fused into the following LayerNorm (one pass instead of two)."""
        # self.attn1(...) goes through the instance attribute so that the VFace hooks, which assign
        # module.forward (ldm/models/pnp_utils.py:289-339), take effect exactly as in the reference.
        x = x.contiguous()
        ln = lambda m, t, **kw: ops.add_layer_norm(t, m.weight, m.bias, m.eps, **kw)
        single = (context is not None and context.shape[1] == 1 and context.shape[-1] != 768 * 2
                  and "forward" not in self.attn2.__dict__)
        if single:
            # attn2's output does not depend on its queries: LN2 is dead and both adds fold into one pass.  bf16: that
            # pass is the to_out GEMM of attn1 itself (x + to_out(a) + b + row in its epilogue, one rounding of the
            # stream); otherwise the LN3 kernel adds a1 and the row while it normalises.
            row = ep = None
            if x.dtype == torch.bfloat16 and _FUSE_TO_OUT:
                row_b = self._row_plus_out_bias(context)
                if row_b is not None:
                    ep = _FusedOut(x, row_b, prebiased=True)
                else:
                    row = self.attn2.single_token_row(context)[:, 0]
                    ep = _FusedOut(x, row)
            self.attn1._fused_out = ep
            try:
                a1 = self._attn1_of_norm1(x, ln)
            finally:
                self.attn1._fused_out = None
            if ep is not None and ep.done:
                x = a1                                      # already x + attn1 + attn2
                n3 = ln(self.norm3, x)
            else:
                if row is None:
                    row = self.attn2.single_token_row(context)[:, 0]
                x, n3 = ln(self.norm3, x, y=a1.contiguous(), row_bias=row)
            return self._ff_residual(x, n3)
        a1 = self._attn1_of_norm1(x, ln)
        x, n2 = ln(self.norm2, x, y=a1.contiguous())
        a2 = self.attn2(n2, context=context)
        x, n3 = ln(self.norm3, x, y=a2.contiguous())
        return self._ff_residual(x, n3)

    def _ff_residual(self, x, n3):
        """x + ff(n3).  bf16 with the reference's GEGLU feed-forward: the down-projection, its bias and the residual add
        are ONE library GEMM (beta = 1), so ff(n3) is never rounded and written out on its own."""
        net = self.ff.net
        if (x.dtype == torch.bfloat16 and isinstance(net[0], GEGLU) and isinstance(net[2], nn.Linear)
                and "forward" not in self.ff.__dict__ and net[2].weight.is_contiguous()):
            h = net[0](n3)                                  # Dropout(p) of an eval-mode network is the identity
            return ops.linear_residual(h.contiguous(), net[2].weight, net[2].bias, x)
        return ops.add_bias(x, self.ff(n3))


class SpatialTransformer(nn.Module):
    """GroupNorm -> 1x1 conv -> tokens -> transformer blocks -> 1x1 conv (zero-init) -> + input
    (reference :246-288).  With channels_last activations 'b c h w -> b (h w) c' is a view."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0., context_dim=None, sep_head_att=False,
                 head_splits=None):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim,
                                   sep_head_att=sep_head_att) for _ in range(depth)])
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))

    def forward(self, x, context=None):
        b, c, h, w = x.shape
        tok = x.permute(0, 2, 3, 1)                              # NHWC view: free for channels_last
        tok = tok.contiguous().view(b, h * w, c)
        g = ops.group_norm_nhwc(tok, self.norm.weight, self.norm.bias, self.norm.eps, self.norm.num_groups)
        w_in = self.proj_in.weight.reshape(self.proj_in.out_channels, c)                                # 1x1 conv
        if ops.linear_proj_supported(g, w_in) and w_in.is_contiguous():
            if _STATS_HANDOFF and _FUSE_LN_QKV and w_in.shape[0] <= 320:
                # the epilogue also emits sum / sum of squares of every output row: the first block's norm1 (folded into its
                # QKV projection) then needs no pass over `t` of its own
                t, st = ops.linear_proj(g, w_in, self.proj_in.bias, emit_stats=True)
                t._vf_row_stats = st
            else:
                t = ops.linear_proj(g, w_in, self.proj_in.bias)
        else:
            t = F.linear(g, w_in, self.proj_in.bias)
        for block in self.transformer_blocks:
            t = block(t, context=context)
        w_out = self.proj_out.weight.reshape(c, -1)                                                     # 1x1 conv
        if ops.linear_proj_supported(t, w_out) and w_out.is_contiguous():
            out = ops.linear_proj(t.contiguous(), w_out, self.proj_out.bias, tok)            # proj_out(t) + x_in, tcgen05 kernel
        elif t.dtype == torch.bfloat16 and w_out.is_contiguous():
            out = ops.linear_residual(t.contiguous(), w_out, self.proj_out.bias, tok)    # proj_out(t) + x_in in one GEMM
        else:
            out = ops.add_bias(F.linear(t, w_out, self.proj_out.bias), tok)
        return out.view(b, h, w, c).permute(0, 3, 1, 2)
