"""Where does DDIMSampler.sample() spend host time outside the 50 denoising steps?  (bench.py's e2e figure sits ~2.5 %
under the device-timed one.)  Times sample() as a whole, the same 50 steps driven directly, and cProfile of sample().

    python experiments/e2e_overhead.py [--frames 32]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=32)
    a = ap.parse_args()
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    device = torch.device("cuda:0")
    model, _ = bench.build_model(device, torch.bfloat16)
    sampler = DDIMSampler(model)
    sampler.make_schedule(50, verbose=False)
    steps = sampler.ddim_timesteps
    clip = bench.local_clip(a.frames, 0, steps)
    host = {k: v.pin_memory() for k, v in clip.items() if isinstance(v, torch.Tensor)}
    host_flow = torch.cat(clip["flow"]).pin_memory()
    host_inv = {t: v.pin_memory() for t, v in clip["inversion"].items()}

    def sample():
        g = lambda t: t.to(device, non_blocking=True)
        inv = {t: g(v) for t, v in host_inv.items()}
        s, _ = sampler.sample(S=50, batch_size=a.frames, shape=(4, 64, 64), conditioning=g(host["c"]),
                              target_conditioning=g(host["target_cond"]), inverse_results_dir=inv, x_T=g(host["x_T"]),
                              flow=g(host_flow), unconditional_guidance_scale=3.0, unconditional_conditioning=g(host["uc"]),
                              eta=0.0, verbose=False,
                              test_model_kwargs=dict(inpaint_image=g(host["inpaint_image"]), inpaint_mask=g(host["inpaint_mask"])))
        torch.cuda.synchronize()
        return s

    sample()
    for _ in range(2):
        t0 = time.perf_counter()
        sample()
        print(f"sample(): {time.perf_counter() - t0:.4f} s")
    # phases
    torch.cuda.synchronize()
    t0 = time.perf_counter(); sampler.make_schedule(50, verbose=False); torch.cuda.synchronize()
    print(f"make_schedule: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    g = lambda t: t.to(device, non_blocking=True)
    t0 = time.perf_counter(); fl = g(host_flow); sampler._register_hooks(fl); torch.cuda.synchronize()
    print(f"_register_hooks: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    t0 = time.perf_counter(); inv = {t: g(v) for t, v in host_inv.items()}; torch.cuda.synchronize()
    print(f"inversion H2D: {(time.perf_counter() - t0) * 1e3:.2f} ms")
    pr = cProfile.Profile()
    pr.enable()
    sample()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(30)


if __name__ == "__main__":
    main()
