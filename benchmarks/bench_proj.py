"""tcgen05 projection kernel (ops.linear_proj) against what it replaces at the 64x64 level of a 32-frame step.

    python benchmarks/bench_proj.py            # CUDA events, L2 flushed before every timed launch
"""
import os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops
from benchmarks.bench_kernels import time_kernel, peaks

pk = peaks()
g = torch.Generator(device="cuda").manual_seed(0)
b, t, k = 96, 4096, 320
rows = b * t
x = (torch.randn(b, t, k, device="cuda", generator=g) + 0.1).bfloat16()
a = torch.randn(b, t, k, device="cuda", generator=g).bfloat16()
ln = torch.nn.LayerNorm(k).cuda().bfloat16()
wq = (torch.randn(960, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
wo = (torch.randn(320, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
bo = torch.randn(320, device="cuda", generator=g).bfloat16()
row = torch.randn(b, 320, device="cuda", generator=g).bfloat16()
unit = rows * 320 * 2 / 1e9      # GB of one (rows, 320) bf16 tensor


def line(name, ms, units):
    gbs = units * unit / (ms * 1e-3)
    print(f"{name:58s} {ms:7.3f} ms  {gbs:7.0f} GB/s  {gbs / pk['hbm']:.2f} of HBM ({units} row-units)")


t_ln, _ = time_kernel(lambda: ops.add_layer_norm(x, ln.weight, ln.bias, ln.eps))
xn = ops.add_layer_norm(x, ln.weight, ln.bias, ln.eps)
t_q, _ = time_kernel(lambda: F.linear(xn, wq))
line("LayerNorm kernel", t_ln, 2)
line("library GEMM 320 -> 960 (QKV)", t_q, 4)
line("  sum: LN + QKV as two launches", t_ln + t_q, 6)
t, _ = time_kernel(lambda: ops.linear_proj(x, wq, None, None, ln=ln))
line("linear_proj LN + QKV, one launch", t, 4)
t, _ = time_kernel(lambda: ops.linear_proj(xn, wq))
line("linear_proj QKV without LN", t, 4)
_, st = ops.linear_proj(a, wo, bo, emit_stats=True)
xs = ops.linear_proj(a, wo, bo)
t, _ = time_kernel(lambda: ops.linear_proj(xs, wq, None, None, ln=ln, ln_stats=st))
line("linear_proj LN + QKV, statistics handed over", t, 4)
t, _ = time_kernel(lambda: ops.linear_proj(a, wo, bo, emit_stats=True))
line("linear_proj proj_in + bias + emitted statistics", t, 2)
rb = (bo.float()[None] + row.float()).bfloat16().contiguous()
t, _ = time_kernel(lambda: ops.linear_residual(a, wo, rb, x))
line("cuBLASLt to_out + per-sample row + residual (batched)", t, 3)
t, _ = time_kernel(lambda: ops.linear_proj(a, wo, bo, x, row_bias=row))
line("linear_proj to_out + per-sample row + residual", t, 3)
t, _ = time_kernel(lambda: ops.linear_residual(a, wo, bo, x))
line("cuBLASLt proj_out + bias + residual", t, 3)
t, _ = time_kernel(lambda: ops.linear_proj(a, wo, bo, x))
line("linear_proj proj_out + bias + residual", t, 3)
t, _ = time_kernel(lambda: F.linear(a, wo, bo))
line("library GEMM proj_in + bias", t, 2)
t, _ = time_kernel(lambda: ops.linear_proj(a, wo, bo))
line("linear_proj proj_in + bias", t, 2)
