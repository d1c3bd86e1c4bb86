"""Correctness + timing of vf_attn_fwd (bf16) under the current VF_ATTN_* environment, one JSON line per case.

    VF_ATTN_STREAM=0|1 VF_ATTN_EMU=0|1|2 python benchmarks/attn_ab.py [--quick]

Correctness: max |got - ref| against an fp32 softmax(q k^T * scale) v on the same bf16-rounded inputs (torch, fp32, on the
GPU), incl. ragged sizes, the second K/V segment, strided q/k/v and a score-jump case.  Timing: CUDA events, L2 flushed.
The environment knobs are read once per process, so A/B runs are separate processes."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops  # noqa: E402
from benchmarks.bench_kernels import time_kernel  # noqa: E402


def ref_attn(q, k, v, heads, k2=None, v2=None):
    b, n, c = q.shape
    d = c // heads
    if k2 is not None:
        k = torch.cat([k, k2], 1)
        v = torch.cat([v, v2], 1)
    sp = lambda t: t.float().reshape(b, -1, heads, d).transpose(1, 2)
    o = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(k), sp(v), scale=d ** -0.5)
    return o.transpose(1, 2).reshape(b, n, c)


def main():
    quick = "--quick" in sys.argv
    env = {k: v for k, v in os.environ.items() if k.startswith("VF_ATTN")}
    torch.backends.cuda.matmul.allow_tf32 = False
    g = torch.Generator(device="cuda").manual_seed(0)
    mk = lambda b, n, c, s=1.0: (torch.randn(b, n, c, device="cuda", generator=g) * s).bfloat16()
    ok = True
    cases = [  # b, nq, nk, heads, d, nk2
        (2, 256, 256, 8, 40, 0), (1, 1024, 1024, 2, 40, 0), (2, 1024, 1024, 8, 80, 0), (2, 256, 256, 8, 160, 0),
        (3, 64, 64, 8, 160, 0), (2, 200, 200, 8, 40, 0), (1, 130, 77, 4, 40, 0), (1, 64, 448, 2, 80, 0),
        (2, 256, 256, 8, 40, 320), (1, 300, 100, 2, 40, 90), (2, 4096, 4096, 8, 40, 0), (1, 128, 64, 1, 128, 0),
        (1, 257, 191, 3, 96, 0), (1, 512, 64, 8, 40, 0), (1, 512, 128, 8, 40, 0),
    ]
    for b, nq, nk, h, d, nk2 in cases:
        q, k, v = mk(b, nq, h * d), mk(b, nk, h * d), mk(b, nk, h * d)
        k2 = mk(b, nk2, h * d) if nk2 else None
        v2 = mk(b, nk2, h * d) if nk2 else None
        got = ops.attention(q, k, v, h, k2=k2, v2=v2).float()
        want = ref_attn(q, k, v, h, k2, v2)
        err = (got - want).abs().max().item()
        good = err < 2e-2 and bool(torch.isfinite(got).all())
        ok &= good
        print(json.dumps(dict(case=f"b{b} nq{nq} nk{nk} h{h} d{d} nk2{nk2}", max_abs_err=err, ok=good, env=env)), flush=True)
    # strided q/k/v out of one buffer + logits that keep growing (lazy rescale) and jump (exact path)
    b, n, h, d = 2, 512, 8, 40
    c = h * d
    qkv = torch.randn(b, n, 3 * c, device="cuda", generator=g)
    qkv[..., :2 * c] *= 3.0
    qkv[:, :, c:2 * c] *= torch.linspace(0.2, 3.0, n, device="cuda")[None, :, None]
    qkv = qkv.bfloat16()
    got = ops.attention(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], h).float()
    want = ref_attn(qkv[..., :c], qkv[..., c:2 * c], qkv[..., 2 * c:], h)
    err = (got - want).abs().max().item()
    good = err < 2e-2 * max(1.0, want.abs().max().item() / 2.0)
    ok &= good
    print(json.dumps(dict(case="strided + growing logits", max_abs_err=err, ok=good, env=env)), flush=True)
    q = torch.randn(2, 448, c, device="cuda", generator=g) * 0.3
    k = torch.randn(2, 448, c, device="cuda", generator=g) * 0.3
    v = torch.randn(2, 448, c, device="cuda", generator=g)
    q[:, 0::3] = 2.0; k[:, 200] = 8.0; k[:, 330] = 1.5; q[:, 1::3] = 1.0; k[:, 401] = 3.0
    q, k, v = q.bfloat16(), k.bfloat16(), v.bfloat16()
    got = ops.attention(q, k, v, h).float()
    want = ref_attn(q, k, v, h)
    err = (got - want).abs().max().item()
    good = err < 2e-2 * max(1.0, want.abs().max().item() / 2.0) and bool(torch.isfinite(got).all())
    ok &= good
    print(json.dumps(dict(case="score jumps (exact path + deferred rescale)", max_abs_err=err, ok=good, env=env)), flush=True)

    # timing
    tcases = [("N=4096 h=8 d=40 frames=96", 96, 4096, 8, 40), ("N=4096 h=8 d=40 frames=16", 16, 4096, 8, 40),
              ("N=1024 h=8 d=80 frames=96", 96, 1024, 8, 80), ("N=256 h=8 d=160 frames=96", 96, 256, 8, 160),
              ("N=64 h=8 d=160 frames=96", 96, 64, 8, 160)]
    if quick:
        tcases = tcases[:1]
    if "--no-timing" in sys.argv:
        tcases = []
    for name, b, n, h, d in tcases:
        q, k, v = mk(b, n, h * d), mk(b, n, h * d), mk(b, n, h * d)
        out = torch.empty_like(q)
        med, best = time_kernel(lambda: ops.attention(q, k, v, h, out=out), iters=10)
        tf = 4.0 * n * n * d * h * b / (med * 1e-3) / 1e12
        print(json.dumps(dict(timing=name, ms=med, ms_best=best, tflops=tf, env=env)), flush=True)
    print("ATTN_AB_OK" if ok else "ATTN_AB_FAIL", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
