"""CTA-pair GEMM+GEGLU (VF_GEMM_PAIR=1, vf_gemm2.cu) against an fp32 reference on the GPU, and its time next to the
single-CTA kernel's.  Run as two processes (the knob is read once):

    VF_GEMM_PAIR=1 python benchmarks/geglu_pair_check.py ; VF_GEMM_PAIR=0 python benchmarks/geglu_pair_check.py
"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops  # noqa: E402
from benchmarks.bench_kernels import time_kernel  # noqa: E402

g = torch.Generator(device="cuda").manual_seed(0)
ok = True
for rows, k, n in ((393216, 320, 1280), (393216 - 100, 320, 1280), (98304, 320, 640), (393216, 256, 1280)):
    x = torch.randn(rows, k, device="cuda", generator=g).bfloat16()
    w = (torch.randn(2 * n, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
    b = (torch.randn(2 * n, device="cuda", generator=g) * 0.1).bfloat16()
    got = ops.linear_geglu(x, w, b)
    torch.cuda.synchronize()
    idx = torch.cat([torch.arange(0, 700, device="cuda"), torch.randint(0, rows, (3000,), device="cuda", generator=g),
                     torch.arange(rows - 700, rows, device="cuda")])
    proj = x[idx].float() @ w.float().t() + b.float()
    val, gate = proj.chunk(2, dim=-1)
    want = val * F.gelu(gate)
    err = (got[idx].float() - want).abs().max().item()
    good = err < 2e-2 * max(1.0, want.abs().max().item() / 4) and bool(torch.isfinite(got).all())
    ok &= good
    med, best = time_kernel(lambda: ops.linear_geglu(x, w, b), 20)
    tf = 2.0 * rows * k * 2 * n / (med * 1e-3) / 1e12
    print(f"VF_GEMM_PAIR={os.environ.get('VF_GEMM_PAIR', '0')} rows={rows} k={k} n={n}: max err {err:.4f} ok={good}  {med:.3f} ms  {tf:.0f} TFLOP/s", flush=True)
print("PAIR_CHECK_OK" if ok else "PAIR_CHECK_FAIL")
sys.exit(0 if ok else 1)
