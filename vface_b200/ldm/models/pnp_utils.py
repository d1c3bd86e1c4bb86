"""VFace attention hooks on the vface_b200 kernels.

Host-side mirror of REFace/ldm/models/pnp_utils.py: find_all_modules_by_name :33-40 and
register_spa_attn_injection :57-339 -- same signature, same plugin mechanism (the `forward`
attribute of every module whose qualified name ends in `attn_component` is replaced by a closure),
same chunk semantics:

    the UNet batch is `chunks` equal parts [uncond ; cond ; recon] (ddim_w_inv.py:654-664); chunk 0
    is always the donor; q and k of chunks 1 and 2 are rewritten in place before attention.

Reachable fusion modes: "replace" :133-143, "fft" :169-183, "flow_fix" :185-222 and
switch_on=False.  The others ("temporal", "adaIn", "mix", "fft_vfixed", chunks == 2) are never
selected by a shipped call site (ddim_w_inv.py:289-308, :389) and raise NotImplementedError.

Execution differs from the reference: one fused projection GEMM, FSAI as one launch per tensor
(both branches share the donor), flow warp as one launch over all frames in the native token layout,
attention without the N x N matrix.  With frame sharding active (vface_b200.frame_shard) the flow
hook exchanges the one-frame halo with the neighbouring ranks.
"""
from __future__ import annotations

import torch

from ... import frame_shard, ops
from ..._autocast import no_autocast

FLOW_TOKENS = 4096          # reference gate `q.shape[1] == 4096` (pnp_utils.py:201): the 64x64 level
FLOW_HW = 64


def find_all_modules_by_name(model, mod_name):
    modules_, names = [], []
    for name, module in model.named_modules():
        if name.endswith(mod_name):
            modules_.append(module)
            names.append(name)
    return modules_, names


def _flow_len(flow):
    if flow is None:
        return 0
    return len(flow) if isinstance(flow, (list, tuple)) else int(flow.shape[0])


def register_spa_attn_injection(model, injection_schedule, switch_on=True, input_blocks=False, output_blocks=True,
                                middle_block=False, attn_component='attn1', chunks=3, flow=None, block_indices=None,
                                fusion="replace", split_ratio_fft=0.8, alpha=0.8, _elide_recon=False):
    """`model` is the sampler (or anything with .model.model.diffusion_model), as in the reference.

    `_elide_recon` is an extension (default off): the batch is [uncond ; cond] only -- the recon
    branch, whose result the reference discards, is not computed; chunk 0 is still the
    donor and only the cond branch is rewritten.
    """
    if fusion not in ("replace", "fft", "flow_fix"):
        raise NotImplementedError(f"fusion='{fusion}' is not reachable from any VFace call site (SURVEY.md 8(a))")
    if chunks != 3:
        raise NotImplementedError("chunks != 3 is not reachable from any VFace call site (SURVEY.md 8(a))")

    def spa_attn_forward(self):

        @no_autocast          # a caller's autocast region would turn the projections into float16 (vface_b200/_autocast.py)
        def forward(x, context=None, mask=None, feature_transfer=True):
            if mask is not None:
                raise NotImplementedError("attention masks are not used on the VFace hot path")
            if context is not None and context is not x:
                # the hooks are only ever registered on self-attention (attn1); keep cross-attention exact
                return type(self).forward(self, x, context=context)
            batch = x.shape[0]
            transfer = feature_transfer and switch_on
            n_parts = 2 if _elide_recon else chunks
            B = batch // n_parts
            q, k, v = self.project_qkv(x)                      # views of one (batch, n, 3C) buffer
            if transfer:
                donor_q, donor_k = q[:B], k[:B]
                cond_q, cond_k = q[B:2 * B], k[B:2 * B]
                rec_q = q[2 * B:] if not _elide_recon else None
                rec_k = k[2 * B:] if not _elide_recon else None
                if fusion == "replace":
                    cond_q.copy_(donor_q)
                    cond_k.copy_(donor_k)
                    if rec_q is not None:
                        rec_q.copy_(donor_q)
                        rec_k.copy_(donor_k)
                else:
                    use_flow = fusion == "flow_fix" and flow is not None and q.shape[1] == FLOW_TOKENS
                    if use_flow:
                        # FSAI writes the cond branch into scratch so the warp (which must not run in
                        # place: frame i+1 reads the un-aligned frame i) lands back in the q/k slices.
                        sq = torch.empty((B, q.shape[1], q.shape[2]), dtype=q.dtype, device=q.device)
                        sk = torch.empty_like(sq)
                    else:
                        sq, sk = cond_q, cond_k
                    if rec_q is not None:
                        ops.fsai_blend2(donor_q, cond_q, rec_q, split_ratio_fft, out_a=sq, out_b=rec_q)
                        ops.fsai_blend2(donor_k, cond_k, rec_k, split_ratio_fft, out_a=sk, out_b=rec_k)
                    else:
                        ops.fsai_blend(donor_q, cond_q, split_ratio_fft, out=sq)
                        ops.fsai_blend(donor_k, cond_k, split_ratio_fft, out=sk)
                    if use_flow:
                        shard = frame_shard.current()
                        sharded = shard is not None and shard.world_size > 1
                        n_flow = _flow_len(flow)
                        want = B if (sharded and shard.rank > 0) else B - 1
                        if n_flow != want:
                            raise ValueError(f"flow has {n_flow} fields for {B} frames"
                                             f"{' + halo' if want == B else ''}; expected {want}")
                        if not sharded:
                            ops.flow_warp_blend(sq, flow, alpha, FLOW_HW, FLOW_HW, out=cond_q)
                            ops.flow_warp_blend(sk, flow, alpha, FLOW_HW, FLOW_HW, out=cond_k)
                        else:
                            # The halo (post-FSAI q/k of the previous rank's last frame) travels on a side stream as soon
                            # as FSAI is done; meanwhile frames 1.. of this shard, which only need their local predecessor,
                            # are warped on the compute stream.  Frame 0 is warped last, against the received halo.
                            pending = shard.exchange_halo_async(sq[B - 1], sk[B - 1])
                            fl = ops._as_flow(flow, sq.device)
                            own = fl[1:] if shard.rank > 0 else fl
                            ops.flow_warp_blend(sq, own, alpha, FLOW_HW, FLOW_HW, out=cond_q)
                            ops.flow_warp_blend(sk, own, alpha, FLOW_HW, FLOW_HW, out=cond_k)
                            halo_q, halo_k = pending.wait()
                            if halo_q is not None:
                                ops.flow_warp_blend(sq[:1], fl[:1], alpha, FLOW_HW, FLOW_HW, prev_halo=halo_q, out=cond_q[:1])
                                ops.flow_warp_blend(sk[:1], fl[:1], alpha, FLOW_HW, FLOW_HW, prev_halo=halo_k, out=cond_k[:1])
            return self.project_out(self.attend(q, k, v))

        forward._vf_native = True      # uses `x` only through project_qkv: BasicTransformerBlock may fold norm1 into it
        return forward

    unet = model.model.model.diffusion_model
    for enabled, blocks in ((input_blocks, unet.input_blocks), (output_blocks, unet.output_blocks),
                            (middle_block, unet.middle_block)):
        if not enabled:
            continue
        mods, _names = find_all_modules_by_name(blocks, attn_component)
        for i, module in enumerate(mods):
            if block_indices is None or i in block_indices:
                module.forward = spa_attn_forward(module)
            else:
                setattr(module, "injection_schedule", injection_schedule)
