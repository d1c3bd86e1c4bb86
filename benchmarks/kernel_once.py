"""A few launches of one kernel at the shapes of a 32-frame step (for ncu):
    python benchmarks/kernel_once.py fsai2|fsai|warp|gn|ln|addln|conv_out|geglu_gemm|linres [bf16|f32] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops
which = sys.argv[1]
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.bfloat16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
frames, n, c = 64, 4096, 320
g = torch.Generator(device="cuda").manual_seed(0)
if which.startswith("fsai"):
    q = torch.randn(3 * frames, n, c, device="cuda", generator=g).to(dt)
    for _ in range(reps):
        if which == "fsai2":
            ops.fsai_blend2(q[:frames], q[frames:2 * frames], q[2 * frames:], 0.8)
        else:
            ops.fsai_blend(q[:frames], q[frames:2 * frames], 0.8, out=q[frames:2 * frames])
elif which in ("gn", "ln", "addln", "conv_out", "geglu_gemm", "linres"):
    b = 96
    rn = lambda *shape: torch.randn(*shape, device="cuda", generator=g).to(dt)
    x, y, w, bias = rn(b, n, c), rn(b, n, c), rn(c), rn(c)
    if which == "conv_out":
        conv = torch.nn.Conv2d(c, 4, 3, padding=1).cuda().to(dt)
        xi = x.view(b, 64, 64, c)
    if which == "geglu_gemm":
        wg, bg = rn(2 * 4 * c, c) / c ** 0.5, rn(2 * 4 * c)
    if which == "linres":
        wd, h = rn(c, 4 * c) / (4 * c) ** 0.5, rn(b, n, 4 * c)
    for _ in range(reps):
        if which == "gn":
            ops.group_norm_nhwc(x, w, bias, 1e-5, 32, silu=True)
        elif which == "ln":
            ops.add_layer_norm(x, w, bias, 1e-5)
        elif which == "addln":
            ops.add_layer_norm(x, w, bias, 1e-5, y=y)
        elif which == "conv_out":
            ops.conv3x3_out_f32(xi, conv)
        elif which == "geglu_gemm":
            ops.linear_geglu(x, wg.contiguous(), bg)
        else:
            ops.linear_residual(h, wd.contiguous(), bias, x)
else:
    x = torch.randn(frames, n, c, device="cuda", generator=g).to(dt)
    flow = (torch.randn(frames - 1, 2, 64, 64, device="cuda", generator=g) * 3).contiguous()
    out = torch.empty_like(x)
    for _ in range(reps):
        ops.flow_warp_blend(x, flow, 0.8, 64, 64, out=out)
torch.cuda.synchronize()
print("ok")
