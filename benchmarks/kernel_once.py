"""One launch of a memory-bound kernel at BASELINE config 4 (for ncu):
    python benchmarks/kernel_once.py fsai2|fsai|warp [bf16|f32] [reps]"""
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
else:
    x = torch.randn(frames, n, c, device="cuda", generator=g).to(dt)
    flow = (torch.randn(frames - 1, 2, 64, 64, device="cuda", generator=g) * 3).contiguous()
    out = torch.empty_like(x)
    for _ in range(reps):
        ops.flow_warp_blend(x, flow, 0.8, 64, 64, out=out)
torch.cuda.synchronize()
print("ok")
