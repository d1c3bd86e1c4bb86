"""Is `residual + x @ W^T (+ bias)` cheaper as ONE library GEMM with beta = 1 than as GEMM + vf_add_bias?"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops
from benchmarks.bench_kernels import time_kernel

g = torch.Generator(device="cuda").manual_seed(0)
for rows, k, n in ((96 * 4096, 1280, 320), (96 * 4096, 320, 320), (96 * 1024, 2560, 640), (96 * 1024, 640, 640), (96 * 256, 5120, 1280), (96 * 256, 1280, 1280)):
    x = torch.randn(rows, k, device="cuda", generator=g).bfloat16()
    w = (torch.randn(n, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
    b = torch.randn(n, device="cuda", generator=g).bfloat16()
    r = torch.randn(rows, n, device="cuda", generator=g).bfloat16()
    wt = w.t()
    t_lin, _ = time_kernel(lambda: F.linear(x, w, b))
    t_two, _ = time_kernel(lambda: ops.add_bias(r, F.linear(x, w, b)))
    rr = r.clone()
    t_inpl, _ = time_kernel(lambda: rr.addmm_(x, wt))
    t_out, _ = time_kernel(lambda: torch.addmm(r, x, wt))
    print(f"rows={rows} k={k} n={n}:  linear+bias {t_lin:.3f} ms | linear+bias, add_bias {t_two:.3f} ms | addmm_ in place (no bias) {t_inpl:.3f} ms | addmm out-of-place {t_out:.3f} ms")

print("per-sample bias through the batched form (to_out + attn2 row + residual) vs flat GEMM + add_layer_norm's add:")
for b, rows, c in ((96, 4096, 320), (96, 1024, 640), (96, 256, 1280)):
    x = torch.randn(b, rows, c, device="cuda", generator=g).bfloat16()
    w = (torch.randn(c, c, device="cuda", generator=g) / c ** 0.5).bfloat16()
    bias = torch.randn(c, device="cuda", generator=g).bfloat16()
    rowb = torch.randn(b, c, device="cuda", generator=g).bfloat16()
    r = torch.randn(b, rows, c, device="cuda", generator=g).bfloat16()
    ln_w, ln_b = torch.ones(c, device="cuda").bfloat16(), torch.zeros(c, device="cuda").bfloat16()
    t_flat, _ = time_kernel(lambda: ops.linear_residual(x, w, bias, r))
    t_bat, _ = time_kernel(lambda: ops.linear_residual(x, w, rowb, r))
    t_lin, _ = time_kernel(lambda: F.linear(x, w, bias))
    t_ln4, _ = time_kernel(lambda: ops.add_layer_norm(r, ln_w, ln_b, 1e-5, y=x, row_bias=rowb))
    t_ln2, _ = time_kernel(lambda: ops.add_layer_norm(r, ln_w, ln_b, 1e-5))
    print(f"b={b} rows={rows} c={c}: flat residual GEMM {t_flat:.3f} | batched per-sample bias {t_bat:.3f} | plain linear {t_lin:.3f} | LN with add (4 units) {t_ln4:.3f} | LN alone (2 units) {t_ln2:.3f}  ->  unfused {t_lin + t_ln4:.3f} vs fused {t_bat + t_ln2:.3f} ms")
