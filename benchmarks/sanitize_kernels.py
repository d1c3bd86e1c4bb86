"""One small launch of every vface_b200 kernel, for compute-sanitizer (memcheck / racecheck / synccheck / initcheck).

    compute-sanitizer --tool memcheck python benchmarks/sanitize_kernels.py

Small shapes (the tools slow a kernel down by 10-100x) that still cover the ragged / strided / second-segment paths.
Prints one line per kernel; exits non-zero if a result is not finite."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops  # noqa: E402


def main():
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, dt=torch.bfloat16: torch.randn(*s, device=dev, generator=g).to(dt)
    ok = True

    def check(name, t):
        nonlocal ok
        torch.cuda.synchronize()
        fin = bool(torch.isfinite(t.float()).all())
        ok &= fin
        print(f"{name:60s} finite={fin}", flush=True)

    # attention: tcgen05 paths (streamed d<=64, split-P d<=160, aliased d 192), ragged sizes, second K/V segment, fp32 path
    for b, nq, nk, h, d, nk2 in ((1, 256, 256, 2, 40, 0), (1, 130, 77, 2, 40, 0), (1, 192, 128, 2, 80, 0), (1, 128, 128, 1, 160, 0),
                                 (1, 128, 64, 1, 192, 0), (1, 128, 128, 2, 40, 90)):
        q, k, v = rn(b, nq, h * d), rn(b, nk, h * d), rn(b, nk, h * d)
        k2 = rn(b, nk2, h * d) if nk2 else None
        v2 = rn(b, nk2, h * d) if nk2 else None
        check(f"attention bf16 nq={nq} nk={nk} h={h} d={d} nk2={nk2}", ops.attention(q, k, v, h, k2=k2, v2=v2))
    q = rn(1, 96, 80, dt=torch.float32)
    check("attention fp32 nq=96 d=40", ops.attention(q, q.clone(), q.clone(), 2))
    # FSAI: the three register-FFT widths (+ an odd row count), the generic Stockham path, both dtypes
    for d in (320, 640, 1280, 160):
        for dt in (torch.bfloat16, torch.float32):
            a, b_, c = rn(1, 37, d, dt=dt), rn(1, 37, d, dt=dt), rn(1, 37, d, dt=dt)
            check(f"fsai_blend d={d} {dt}", ops.fsai_blend(a, b_, 0.8))
            oa, ob = ops.fsai_blend2(a, b_.clone(), c.clone(), 0.8)
            check(f"fsai_blend2 d={d} {dt}", oa + ob)
    # flow warp (+ halo), both dtypes
    for dt in (torch.bfloat16, torch.float32):
        x = rn(3, 256, 64, dt=dt)
        fl = torch.randn(2, 2, 16, 16, device=dev, generator=g) * 3
        check(f"flow_warp_blend {dt}", ops.flow_warp_blend(x, fl, 0.8, 16, 16))
        fl3 = torch.randn(3, 2, 16, 16, device=dev, generator=g) * 30
        check(f"flow_warp_blend + halo {dt}", ops.flow_warp_blend(x, fl3, 0.8, 16, 16, prev_halo=x[0].clone()))
    # CFG + DDIM, inversion step
    x = rn(2, 4, 16, 16, dt=torch.float32)
    for dt in (torch.float32, torch.bfloat16):
        eu, ec = rn(2, 4, 16, 16, dt=dt), rn(2, 4, 16, 16, dt=dt)
        xp, p0 = ops.ddim_cfg_step(x, eu, ec, 0.5, 0.6, 0.1, 0.7, 3.0, noise=rn(2, 4, 16, 16, dt=torch.float32))
        check(f"ddim_cfg_step e={dt}", xp + p0)
        check(f"ddim_invert_step e={dt}", ops.ddim_invert_step(x, ec, 0.5, 0.45, e_uncond=eu, cfg_scale=2.0))
    # glue: GroupNorm (+cat, +add, +SiLU), LayerNorm (+add, +row bias), GEGLU, add_bias, upsample, output convolution
    for dt in (torch.bfloat16, torch.float32):
        t = rn(2, 64, 320, dt=dt)
        w, bb = rn(320, dt=dt), rn(320, dt=dt)
        check(f"group_norm {dt}", ops.group_norm_nhwc(t, w, bb, 1e-5, 32, silu=True, add_nc=rn(2, 320, dt=dt)))
        w2, b2 = rn(640, dt=dt), rn(640, dt=dt)
        check(f"group_norm cat {dt}", ops.group_norm_nhwc(t, w2, b2, 1e-5, 32, silu=True, x2=rn(2, 64, 320, dt=dt)))
        check(f"layer_norm {dt}", ops.add_layer_norm(t, w, bb))
        r, o = ops.add_layer_norm(t, w, bb, y=rn(2, 64, 320, dt=dt), row_bias=rn(2, 320, dt=dt))
        check(f"add_layer_norm {dt}", r + o)
        check(f"geglu {dt}", ops.geglu(rn(2, 64, 640, dt=dt)))
        check(f"add_bias {dt}", ops.add_bias(t, rn(2, 64, 320, dt=dt), rn(2, 320, dt=dt)))
        xi = rn(1, 64, 8, 8, dt=dt).contiguous(memory_format=torch.channels_last)
        check(f"upsample_nearest2x {dt}", ops.upsample_nearest2x(xi))
    conv = torch.nn.Conv2d(320, 4, 3, padding=1).to(dev).to(torch.bfloat16)
    check("conv3x3_out_f32", ops.conv3x3_out_f32(rn(1, 16, 16, 320), conv))
    # fused tcgen05 GEMM + GEGLU: W-resident (k = 320) and streaming (k = 640) kernels
    for k, n in ((320, 256), (640, 256)):
        xg, wg, bg = rn(600, k), rn(2 * n, k) * 0.05, rn(2 * n)
        check(f"linear_geglu k={k} n={n}", ops.linear_geglu(xg, wg, bg))
    print("SANITIZE_RUN_OK" if ok else "SANITIZE_RUN_NONFINITE", flush=True)
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
