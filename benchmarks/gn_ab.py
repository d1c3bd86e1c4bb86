"""GroupNorm A/B: the fused persistent kernel (one launch, apply pass out of L2) against the two-launch form.

    python benchmarks/gn_ab.py                 # runs every configuration in its own process, prints one table
    python benchmarks/gn_ab.py --child         # one configuration (environment decides), JSON on stdout

Shapes are those of one 32-frame step (UNet batch 96).  Timing as in bench_kernels.py: CUDA events, 3 warm-up
launches, a 256 MiB L2 flush before every timed launch; bytes = 2 reads + 1 write of x (the algorithmic figure
of DESIGN.md 3.5, also for the fused kernel, whose second read is an L2 hit).  Each child also checks its output
against torch.nn.functional.group_norm in fp32.
"""
from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SHAPES = [            # (n, hw, c1, c2, silu, add)
    (96, 4096, 320, 0, True, True),
    (96, 4096, 320, 0, False, False),
    (96, 1024, 640, 0, True, True),
    (96, 256, 1280, 0, True, True),
    (96, 64, 1280, 0, True, False),
    (96, 4096, 320, 320, True, False),
    (96, 4096, 640, 320, True, False),
    (96, 1024, 1280, 640, True, False),
    (96, 256, 1280, 1280, True, False),
    (6, 4096, 320, 0, True, True),
]

CONFIGS = [
    ("two launches", dict(VF_GN_FUSED="0")),
    ("fused, L2 re-read, 160K", dict(VF_GN_FUSED="1", VF_GN_FUSED_CTAS="2", VF_GN_SLAB_KB="160")),
    ("slab in smem, fewest slabs", dict(VF_GN_FUSED="2", VF_GN_RES_QUANT="0")),
    ("slab in smem, wave plan", dict(VF_GN_FUSED="2", VF_GN_RES_QUANT="1")),
]


def child():
    import torch
    import torch.nn.functional as F
    from benchmarks.bench_kernels import time_kernel
    from vface_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    out = []
    for n, hw, c1, c2, silu, add in SHAPES:
        c = c1 + c2
        x1 = (torch.randn(n, hw, c1, device="cuda", generator=g) * 1.5 + 0.3).bfloat16()
        x2 = (torch.randn(n, hw, c2, device="cuda", generator=g) - 0.2).bfloat16() if c2 else None
        w = torch.randn(c, device="cuda", generator=g).bfloat16()
        b = torch.randn(c, device="cuda", generator=g).bfloat16()
        a = torch.randn(n, c, device="cuda", generator=g).bfloat16() if add else None
        fn = lambda: ops.group_norm_nhwc(x1, w, b, 1e-5, 32, silu=silu, add_nc=a, x2=x2)
        y = fn()
        xf = (torch.cat([x1, x2], -1) if c2 else x1).float()
        if add:
            xf = xf + a.float()[:, None, :]
        ref = F.group_norm(xf.transpose(1, 2), 32, w.float(), b.float(), 1e-5).transpose(1, 2)
        if silu:
            ref = F.silu(ref)
        err = ((y.float() - ref).abs() / (ref.abs() + 1.0)).max().item()       # bf16 rounding of the output: <= 2^-8 relative
        y2 = fn()
        same = bool(torch.equal(y, y2))
        med, best = time_kernel(fn, 15)
        by = 3.0 * n * hw * c * 2
        out.append(dict(shape=[n, hw, c1, c2, int(silu), int(add)], ms=med, ms_best=best, gbs=by / (med * 1e-3) / 1e9, err=err, rerun_equal=same))
    print("RESULT " + json.dumps(out))


def main():
    if "--child" in sys.argv:
        return child()
    rows = {}
    for name, env in CONFIGS:
        e = dict(os.environ)
        e.update(env)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child"], env=e, capture_output=True, text=True, timeout=900)
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")]
        if not line:
            print(f"{name}: FAILED\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
            continue
        rows[name] = json.loads(line[-1][7:])
    peak = 6557.1
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    print(f"{'shape (n, hw, c1, c2, silu, add)':38s} " + " | ".join(f"{k:>26s}" for k in rows))
    for i, sh in enumerate(SHAPES):
        cells = []
        for k in rows:
            d = rows[k][i]
            flag = "" if (d["err"] < 1.2e-2 and d["rerun_equal"]) else " !!"
            cells.append(f"{d['ms']:7.3f} ms {100 * d['gbs'] / peak:5.1f}% e={d['err']:.0e}{flag}")
        print(f"{str(sh):38s} " + " | ".join(f"{c:>26s}" for c in cells))


if __name__ == "__main__":
    main()
