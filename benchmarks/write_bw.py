"""How fast can this GPU WRITE?  The copy peak (6.56 TB/s) is half reads, half writes; the QKV projection of the 64x64 level
reads 1 row-unit and writes 3.  Times a pure fill, a 1-read-3-write expansion (the QKV traffic mix) and a plain copy."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from benchmarks.bench_kernels import time_kernel
rows, c = 96 * 4096, 320
x = torch.randn(rows, c, device="cuda").bfloat16()
y3 = torch.empty(rows, 3 * c, device="cuda", dtype=torch.bfloat16)
y1 = torch.empty(rows, c, device="cuda", dtype=torch.bfloat16)
unit = rows * c * 2 / 1e9
for name, fn, units in (("fill 3 units (pure write)", lambda: y3.zero_(), 3),
                        ("expand 1 read -> 3 written (QKV mix)", lambda: torch.cat([x, x, x], dim=1, out=y3), 4),
                        ("copy 1 read -> 1 written", lambda: y1.copy_(x), 2),
                        ("copy 3 -> 3", lambda: y3.copy_(y3.roll(1, 0)) if False else y3.copy_(y3_src), 6)):
    if "3 -> 3" in name:
        y3_src = torch.randn(rows, 3 * c, device="cuda").bfloat16()
    t, tmin = time_kernel(fn)
    print(f"{name:44s} {t:7.3f} ms (min {tmin:.3f})  {units * unit / t * 1e3:7.0f} GB/s")
