"""Where does the bf16 path's error come from?  One 3-branch UNet call at full size in fp32 and in bf16 (same weights,
same inputs); relative L2 error of every top-level block's output, of the final eps per branch and of the CFG
combination e_u + 3 (e_c - e_u)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.test_pipeline_gpu import build, rel_l2
from vface_b200 import synth

clip = synth.synth_clip(2)
x = torch.cat([clip["x_T"], clip["inpaint_image"], clip["inpaint_mask"]], dim=1)
x_in = torch.cat([x, x, torch.cat([torch.randn(2, 4, 64, 64, generator=torch.Generator().manual_seed(3)), x[:, 4:]], 1)]).cuda()
ctx = torch.cat([clip["uc"], clip["c"], clip["target_cond"]]).cuda()
t = torch.full((6,), 801, dtype=torch.long, device="cuda")
acts = {}
for dtype in (torch.float32, torch.bfloat16):
    model, _, _ = build(None, dtype)
    unet = model.model.diffusion_model
    rec = {}
    hooks = []
    def mk(name):
        def hook(m, i, o):
            o = o[0] if isinstance(o, tuple) else o
            rec[name] = o.detach().float().cpu()
        return hook
    for i, m in enumerate(unet.input_blocks): hooks.append(m.register_forward_hook(mk(f"in{i:02d}")))
    hooks.append(unet.middle_block.register_forward_hook(mk("mid")))
    for i, m in enumerate(unet.output_blocks): hooks.append(m.register_forward_hook(mk(f"out{i:02d}")))
    with torch.no_grad():
        rec["eps"] = model.apply_model(x_in, t, ctx).float().cpu()
    acts[dtype] = rec
    del model
    torch.cuda.empty_cache()
a, b = acts[torch.float32], acts[torch.bfloat16]
for k in a:
    if k == "eps": continue
    d_ref = a[k][2:4] - a[k][0:2]          # cond - uncond
    d_got = b[k][2:4] - b[k][0:2]
    print(f"{k:6s} shape {tuple(a[k].shape)}  rel err all {rel_l2(b[k], a[k]):.5f}   |cond-uncond|/|act| {float(d_ref.norm() / a[k][0:2].norm()):.4f}   rel err of (cond-uncond) {rel_l2(d_got, d_ref):.5f}")
e, g = a["eps"], b["eps"]
for i, n in enumerate(("uncond", "cond", "recon")):
    print(f"eps {n:6s} rel err {rel_l2(g[2 * i:2 * i + 2], e[2 * i:2 * i + 2]):.5f}  norm {float(e[2 * i:2 * i + 2].norm()):.2f}")
cfg = lambda z: z[0:2] + 3.0 * (z[2:4] - z[0:2])
print(f"eps cfg    rel err {rel_l2(cfg(g), cfg(e)):.5f}  norm {float(cfg(e).norm()):.2f};  |e_c - e_u| = {float((e[2:4] - e[0:2]).norm()):.2f}, rel err of the difference {rel_l2(g[2:4] - g[0:2], e[2:4] - e[0:2]):.5f}")
