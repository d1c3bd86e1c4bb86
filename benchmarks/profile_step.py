"""Kernel-time breakdown of one denoising step (bench.py workload) with torch.profiler.

    python benchmarks/profile_step.py [--frames 32] [--top 40]

Not a benchmark (profiler overhead): use it for SHARES of the step, like the ncu launch list.
"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--elide-recon", action="store_true")
    ap.add_argument("--shapes", default="", help="comma list of aten op names to break down by input shape (e.g. aten::add,aten::copy_)")
    a = ap.parse_args()
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    device = torch.device("cuda:0")
    model, _ = bench.build_model(device, torch.bfloat16)
    sampler = DDIMSampler(model, elide_dead_recon=a.elide_recon)
    sampler.make_schedule(50, verbose=False)
    steps = sampler.ddim_timesteps
    clip = bench.local_clip(a.frames, 0, steps)
    dev = {k: v.to(device) for k, v in clip.items() if isinstance(v, torch.Tensor)}
    flow = torch.cat(clip["flow"]).to(device)
    inv = {t: v.to(device) for t, v in clip["inversion"].items()}
    sampler._register_hooks(flow)
    sampler._inv_cache = inv
    kw = dict(test_model_kwargs=dict(inpaint_image=dev["inpaint_image"], inpaint_mask=dev["inpaint_mask"]))
    tr = np.flip(steps)

    def one(i, x):
        step = int(tr[i])
        ts = torch.full((a.frames,), step, device=device, dtype=torch.long)
        return sampler.p_sample_ddim_with_inverse(x, dev["c"], ts, index=49 - i, target_conditioning=dev["target_cond"],
                                                  inverse_results_dir=inv, unconditional_guidance_scale=3.0,
                                                  unconditional_conditioning=dev["uc"], flow=flow, _step=step, **kw)[0]

    x = dev["x_T"]
    with torch.no_grad():
        for i in range(3):
            x = one(i, x)
        torch.cuda.synchronize()
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], record_shapes=bool(a.shapes)) as prof:
            for i in range(3, 5):
                x = one(i, x)
            torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=a.top, max_name_column_width=90))
    if a.shapes:
        want = set(a.shapes.split(","))
        rows = [e for e in prof.key_averages(group_by_input_shape=True) if e.key in want]
        rows.sort(key=lambda e: -e.device_time_total)
        for e in rows[:40]:
            print(f"{e.key:18s} calls={e.count:4d} cuda_total={e.device_time_total / 1e3:8.3f} ms  shapes={e.input_shapes}")


if __name__ == "__main__":
    main()
