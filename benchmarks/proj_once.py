"""One launch each of the projection kernel's four forms at the size of a 32-frame step (for ncu): LN + QKV, to_out + row + residual."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
b, t, k = 96, 4096, 320
x = (torch.randn(b, t, k, device="cuda", generator=g) + 0.1).bfloat16()
a = torch.randn(b, t, k, device="cuda", generator=g).bfloat16()
ln = torch.nn.LayerNorm(k).cuda().bfloat16()
wq = (torch.randn(960, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
wo = (torch.randn(320, k, device="cuda", generator=g) / k ** 0.5).bfloat16()
bo = torch.randn(320, device="cuda", generator=g).bfloat16()
row = torch.randn(b, 320, device="cuda", generator=g).bfloat16()
ops._ProjFold.get(wq, None, ln); ops._ProjFold.get(wo, bo, None)
torch.cuda.synchronize()
t, st = ops.linear_proj(a, wo, bo, emit_stats=True)          # proj_in form: emits the row statistics
ops.linear_proj(t, wq, None, None, ln=ln, ln_stats=st)        # LN + QKV with the statistics handed over
ops.linear_proj(x, wq, None, None, ln=ln)                     # LN + QKV with in-kernel statistics
ops.linear_proj(a, wo, bo, x, row_bias=row)                   # to_out + per-sample row + residual
torch.cuda.synchronize()
print("ok")
