"""Kernel microbenchmarks: BASELINE.json configs 3 (attention) and 4 (FSAI, flow warp) + CFG/DDIM.

    python benchmarks/bench_kernels.py [--only attn|fsai|warp|ddim|glue|gemm] [--iters 20] [--json out.json]

Timing: CUDA events on the launching (current) stream, >= 3 warm-up launches, an L2 flush (write of a
256 MiB buffer) before every timed launch.  Rooflines use /root/repo/MEASURED_PEAKS.json when present,
else the fallback of B200_PROFILING.md (6.65 TB/s, 1.59 PFLOP/s) -- the output says which.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vface_b200 import ops  # noqa: E402


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc=d["bf16_tflops"], tc_sustained=d.get("bf16_tflops_sustained"), src="measured")
    return dict(hbm=6650.0, tc=1590.0, tc_sustained=1400.0, src="fallback")


_flush = None


def time_kernel(fn, iters=20, warmup=3):
    global _flush
    if _flush is None:
        _flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        _flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def bench_attn(res, iters):
    pk = peaks()
    cases = [
        ("attn N=4096 h=8 d=40 frames=16", 16, 4096, 4096, 8, 40, 0),
        ("attn N=4096 h=8 d=40 frames=16 +concat KV (Nk=8192)", 16, 4096, 4096, 8, 40, 4096),
        ("attn N=4096 h=8 d=40 frames=96 (32 frames x 3 branches)", 96, 4096, 4096, 8, 40, 0),
        ("attn N=1024 h=8 d=80 frames=96", 96, 1024, 1024, 8, 80, 0),
        ("attn N=256 h=8 d=160 frames=96", 96, 256, 256, 8, 160, 0),
        ("attn N=64 h=8 d=160 frames=96", 96, 64, 64, 8, 160, 0),
    ]
    for name, b, nq, nk, h, d, nk2 in cases:
        g = torch.Generator(device="cuda").manual_seed(0)
        mk = lambda n: torch.randn(b, n, h * d, device="cuda", generator=g).bfloat16()
        q, k, v = mk(nq), mk(nk), mk(nk)
        k2 = mk(nk2) if nk2 else None
        v2 = mk(nk2) if nk2 else None
        out = torch.empty_like(q)
        fn = lambda: ops.attention(q, k, v, h, k2=k2, v2=v2, out=out)
        med, best = time_kernel(fn, iters)
        flops = 4.0 * nq * (nk + nk2) * d * h * b
        tf = flops / (med * 1e-3) / 1e12
        res.append(dict(kernel=name, ms=med, ms_best=best, tflops=tf, frac_burst=tf / pk["tc"],
                        frac_sustained=tf / pk["tc_sustained"] if pk["tc_sustained"] else None, bound="tensor", peaks=pk["src"]))
        print(f"{name:62s} {med:8.3f} ms  {tf:8.1f} TFLOP/s  {100 * tf / pk['tc']:5.1f}% of {pk['src']} burst peak")


def bench_fsai(res, iters):
    pk = peaks()
    for dt, e in ((torch.bfloat16, 2), (torch.float32, 4)):
        for n, c in ((4096, 320), (1024, 640), (256, 1280)):
            frames = 64
            g = torch.Generator(device="cuda").manual_seed(0)
            q = torch.randn(3 * frames, n, c, device="cuda", generator=g).to(dt)
            fn1 = lambda: ops.fsai_blend(q[:frames], q[frames:2 * frames], 0.8, out=q[frames:2 * frames])
            med, best = time_kernel(fn1, iters)
            by = 3.0 * n * c * e * frames
            gbs = by / (med * 1e-3) / 1e9
            name = f"fsai single {dt} N={n} C={c} frames={frames}"
            res.append(dict(kernel=name, ms=med, ms_best=best, gbs=gbs, frac=gbs / pk["hbm"], bound="hbm", peaks=pk["src"]))
            print(f"{name:62s} {med:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / pk['hbm']:5.1f}% of {pk['src']} HBM")
            fn2 = lambda: ops.fsai_blend2(q[:frames], q[frames:2 * frames], q[2 * frames:], 0.8)
            med, best = time_kernel(fn2, iters)
            by = 5.0 * n * c * e * frames
            gbs = by / (med * 1e-3) / 1e9
            name = f"fsai fused2 {dt} N={n} C={c} frames={frames}"
            res.append(dict(kernel=name, ms=med, ms_best=best, gbs=gbs, frac=gbs / pk["hbm"], bound="hbm", peaks=pk["src"]))
            print(f"{name:62s} {med:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / pk['hbm']:5.1f}% of {pk['src']} HBM")


def bench_warp(res, iters):
    pk = peaks()
    for dt, e in ((torch.bfloat16, 2), (torch.float32, 4)):
        for hw, c in ((64, 320), (32, 640)):
            frames = 64
            n = hw * hw
            g = torch.Generator(device="cuda").manual_seed(0)
            x = torch.randn(frames, n, c, device="cuda", generator=g).to(dt)
            flow = (torch.randn(frames - 1, 2, hw, hw, device="cuda", generator=g) * 3).contiguous()
            out = torch.empty_like(x)
            fn = lambda: ops.flow_warp_blend(x, flow, 0.8, hw, hw, out=out)
            med, best = time_kernel(fn, iters)
            by = (3.0 * n * c * e + 8.0 * n) * (frames - 1) + 2.0 * n * c * e
            gbs = by / (med * 1e-3) / 1e9
            name = f"flow warp+blend {dt} {hw}x{hw} C={c} frames={frames}"
            res.append(dict(kernel=name, ms=med, ms_best=best, gbs=gbs, frac=gbs / pk["hbm"], bound="hbm", peaks=pk["src"]))
            print(f"{name:62s} {med:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / pk['hbm']:5.1f}% of {pk['src']} HBM")


def bench_ddim(res, iters):
    pk = peaks()
    for frames in (32, 256, 4096):
        g = torch.Generator(device="cuda").manual_seed(0)
        mk = lambda: torch.randn(frames, 4, 64, 64, device="cuda", generator=g)
        x, eu, ec = mk(), mk(), mk()
        fn = lambda: ops.ddim_cfg_step(x, eu, ec, 0.5, 0.6, 0.0, 0.7071, 3.0)
        med, best = time_kernel(fn, iters)
        by = 5.0 * frames * 16384 * 4
        gbs = by / (med * 1e-3) / 1e9
        name = f"cfg+ddim fp32 frames={frames}"
        res.append(dict(kernel=name, ms=med, ms_best=best, gbs=gbs, frac=gbs / pk["hbm"], bound="hbm", peaks=pk["src"]))
        print(f"{name:62s} {med:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / pk['hbm']:5.1f}% of {pk['src']} HBM")


def bench_glue(res, iters):
    """The fused glue kernels of the UNet at the shapes of one 32-frame step (UNet batch 96)."""
    pk = peaks()
    dt, e = torch.bfloat16, 2
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g).to(dt)

    def rec(name, fn, by):
        med, best = time_kernel(fn, iters)
        gbs = by / (med * 1e-3) / 1e9
        res.append(dict(kernel=name, ms=med, ms_best=best, gbs=gbs, frac=gbs / pk["hbm"], bound="hbm", peaks=pk["src"]))
        print(f"{name:62s} {med:8.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / pk['hbm']:5.1f}% of {pk['src']} HBM")

    for n, c in ((4096, 320), (1024, 640), (256, 1280)):
        b = 96
        h = rn(b, n, 8 * c)
        rec(f"geglu rows={b * n} k={4 * c}", lambda: ops.geglu(h), 3.0 * b * n * 4 * c * e)
        x = rn(b, n, c)
        w, bb = rn(c), rn(c)
        rec(f"group_norm+silu NHWC n={b} hw={n} c={c} (2 reads + 1 write)", lambda: ops.group_norm_nhwc(x, w, bb, 1e-5, 32, silu=True),
            3.0 * b * n * c * e)
        y = rn(b, n, c)
        rec(f"add_bias rows={b * n} c={c}", lambda: ops.add_bias(x, y, bb), 3.0 * b * n * c * e)
        rec(f"add+layer_norm rows={b * n} c={c} (2 reads + 2 writes)", lambda: ops.add_layer_norm(x, w, bb, 1e-5, y=y), 4.0 * b * n * c * e)
        rec(f"layer_norm rows={b * n} c={c} (1 read + 1 write)", lambda: ops.add_layer_norm(x, w, bb, 1e-5), 2.0 * b * n * c * e)
    # output convolution with an fp32 result (9 c 4 MAC per pixel; reads x once, writes 4 fp32 per pixel)
    conv = torch.nn.Conv2d(320, 4, 3, padding=1).cuda().bfloat16()
    xo = rn(96, 64, 64, 320)
    rec("conv3x3 320->4, fp32 out, n=96 64x64 (1 read of x)", lambda: ops.conv3x3_out_f32(xo, conv), 96 * 4096 * (320 * e + 16.0))
    x1, x2 = rn(96, 4096, 320), rn(96, 4096, 320)
    w, bb = rn(640), rn(640)
    rec("group_norm+silu over cat(320+320) n=96 hw=4096", lambda: ops.group_norm_nhwc(x1, w, bb, 1e-5, 32, silu=True, x2=x2),
        3.0 * 96 * 4096 * 640 * e)


def bench_gemm(res, iters):
    """Fused tcgen05 GEMM+GEGLU against library GEMM + geglu kernel at the three feed-forward shapes of a step."""
    import torch.nn.functional as F
    pk = peaks()
    g = torch.Generator(device="cuda").manual_seed(0)
    for rows, k, n in ((96 * 4096, 320, 1280), (96 * 1024, 640, 2560), (96 * 256, 1280, 5120)):
        x = torch.randn(rows, k, device="cuda", generator=g).bfloat16()
        w = (torch.randn(2 * n, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
        b = torch.randn(2 * n, device="cuda", generator=g).bfloat16()
        flops = 2.0 * rows * k * 2 * n
        for name, fn in (("fused linear+geglu (tcgen05)", lambda: ops.linear_geglu(x, w, b)),
                         ("cuBLAS linear + vf_geglu", lambda: ops.geglu(F.linear(x, w, b))),
                         ("cuBLAS linear alone", lambda: F.linear(x, w, b))):
            med, best = time_kernel(fn, iters)
            tf = flops / (med * 1e-3) / 1e12
            full = f"{name} rows={rows} k={k} n={n}"
            res.append(dict(kernel=full, ms=med, ms_best=best, tflops=tf, frac_burst=tf / pk["tc"], bound="tensor", peaks=pk["src"]))
            print(f"{full:62s} {med:8.3f} ms  {tf:8.1f} TFLOP/s  {100 * tf / pk['tc']:5.1f}% of {pk['src']} burst peak")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--json", default="")
    a = ap.parse_args()
    res = []
    print(torch.cuda.get_device_name(0), "| peaks:", peaks())
    for name, fn in (("attn", bench_attn), ("fsai", bench_fsai), ("warp", bench_warp), ("ddim", bench_ddim), ("glue", bench_glue), ("gemm", bench_gemm)):
        if a.only and a.only != name:
            continue
        fn(res, a.iters)
    if a.json:
        os.makedirs(os.path.dirname(os.path.abspath(a.json)), exist_ok=True)
        json.dump(res, open(a.json, "w"), indent=1)


if __name__ == "__main__":
    main()
