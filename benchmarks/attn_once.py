"""One attention launch at BASELINE config 3 (for ncu): python benchmarks/attn_once.py [frames] [reps]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vface_b200 import ops
frames = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator(device="cuda").manual_seed(0)
mk = lambda: torch.randn(frames, 4096, 320, device="cuda", generator=g).bfloat16()
q, k, v = mk(), mk(), mk()
out = torch.empty_like(q)
for _ in range(reps):
    ops.attention(q, k, v, 8, out=out)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
