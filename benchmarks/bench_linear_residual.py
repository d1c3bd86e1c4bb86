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
