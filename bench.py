#!/usr/bin/env python
"""Headline benchmark of the VFace denoising hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--frames F] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

Workload (BASELINE.json configs[1]): full VFace pipeline, 32-frame 512x512 clip per GPU, DDIM-50 with
CFG 3.0, bf16 kernels, random-init weights of the named UNet (859.5 M parameters, every zero-module
re-randomised) and synthetic clips -- no checkpoints or datasets exist offline.

  step      one denoising step (DDIMSampler.p_sample_ddim_with_inverse: 3-way UNet batch with the
            VFace hooks on + fused CFG/DDIM update) over the rank's 32 frames; consecutive timed steps
            walk down the DDIM-50 schedule.
  value     frames/s at DDIM-50 = total frames / (50 * step time); inputs resident in HBM.
  e2e       the same metric through the public API, DDIMSampler.sample(S=50) on HOST (pinned)
            buffers: all host->device copies (clip, conditioning, flow, 50 inversion latents) and the
            device->host read of the samples are inside the timed region.
  roofline  the dominant kernel (fused tcgen05 attention at N=4096, 8 heads x d40), timed live with
            CUDA events around every launch inside the timed steps.
  N > 1     frames shard by contiguous chunk (weak scaling: 32 frames per GPU, so N=8 is the 256-frame
            clip); the only exchange is the one-frame q/k halo per flow-active module per step.

--impl reference times the reference's CPU path (oracle port, kind "port": the reference is Python and
cannot travel to the GPU box) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

DDIM_STEPS = 50
CFG_SCALE = 3.0
METRIC = "frames/s at 512^2 DDIM-50 (VFace denoising hot path)"
UNIT = "frames/s"
# attn1 modules at N=4096 (64x64, 8 heads x d40): input_blocks.1/.2, output_blocks.9/.10/.11
ATTN4096_FLOPS_PER_SAMPLE = 4.0 * 4096 * 4096 * 40 * 8


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tc_burst=d["bf16_tflops"], tc_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), src="measured")
    return dict(hbm=6650.0, tc_burst=1590.0, tc_sustained=1400.0, src="fallback")


# ---- clocks -----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.nv = None

    def run(self):
        if not self.ok:
            return
        while not self._stop_evt.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                try:
                    mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        med = float(np.median(self.samples)) if self.samples else None
        return dict(sm_mhz=med, sm_max_mhz=self.max_mhz, reasons=sorted(self.reasons), samples=len(self.samples))


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ---- model / clip ---------------------------------------------------------------------------------------
def build_model(device, dtype, state=None):
    from vface_b200 import synth
    from vface_b200.latent_diffusion import LatentDiffusion
    model = LatentDiffusion()
    unet = model.model.diffusion_model
    sd = state if state is not None else synth.synth_state_dict(unet.state_dict(), seed=1)
    unet.load_state_dict(sd)
    model = model.to(device).eval()
    unet.to(dtype)
    return model, sd


def local_clip(frames, rank, steps):
    from vface_b200 import synth
    clip = synth.synth_clip(frames + (1 if rank > 0 else 0), seed=7 + 101 * rank, steps=steps)
    if rank > 0:
        # one extra leading flow field (halo frame -> first local frame); per-frame tensors drop the extra frame
        for k in ("x_T", "inpaint_image", "inpaint_mask", "c", "target_cond", "uc"):
            clip[k] = clip[k][1:].contiguous()
        clip["inversion"] = {t: v[1:].contiguous() for t, v in clip["inversion"].items()}
    return clip


CPU_ARM_FRAMES = 2      # frames per step the CPU arm (oracle port) actually runs


def workload_config(frames, world, n_branches=3, strong=False):
    """The `config` object of the JSON line -- identical for the vface_b200 arm and the reference arm."""
    workload = "VFace full pipeline, 32-frame 512x512 clip per GPU, DDIM-50 with CFG 3.0, bf16"
    if strong:      # BASELINE.json configs[4]
        workload = (f"{frames * world}-frame 512x512 clip sharded by frame across {world} B200 with NCCL halo exchange of boundary "
                    f"attention features, DDIM-50 with CFG 3.0, bf16")
    return dict(workload=workload,
                scope="denoising loop only: 50 x (3-branch UNet forward with the VFace hooks + CFG/DDIM update); VAE, DDIM "
                      "inversion and flow estimation run once per clip outside the loop and are not timed",
                frames_per_gpu=frames, total_frames=frames * world, ddim_steps=DDIM_STEPS, cfg_scale=CFG_SCALE,
                unet_batch_per_step=n_branches * frames, branches=n_branches, parallelism=f"frame-shard x{world}",
                l2="step working set (1.7 GB weights + activations) far exceeds the 126 MB L2; no flush needed",
                weights="random-init REFace UNet 859.5M params, zero-modules re-randomised (seed 1)",
                cpu_arm_sample=f"the CPU arm (--impl reference, and cpu_baseline) times a bounded sample of this workload: "
                               f"{CPU_ARM_FRAMES} of the frames per step (UNet batch {3 * CPU_ARM_FRAMES}, hooks and flow warp on), "
                               f"one host process, frames/s normalised per frame")


# ---- CPU baseline (oracle port) -------------------------------------------------------------------------------
def cpu_step_seconds(sd, frames, warm, reps):
    """Seconds per denoising step of the reference path on the host cores (oracle port), `frames` frames."""
    from oracle import kernels as ok, port
    from vface_b200 import synth
    steps = ok.make_schedule(DDIM_STEPS)["ddim_timesteps"]
    clip = synth.synth_clip(max(frames, 2), steps=steps[-(warm + reps):])
    take = lambda t: t[:frames]
    ts = []
    hooks = port.vface_hooks(sd, clip["flow"][:frames - 1] if frames > 1 else None)
    extra = torch.cat([take(clip["inpaint_image"]), take(clip["inpaint_mask"])], dim=1)
    x = take(clip["x_T"])
    for i, step in enumerate(np.flip(steps)[:warm + reps]):
        t0 = time.perf_counter()
        tt = torch.full((frames,), int(step), dtype=torch.long)
        x_full = torch.cat([x, extra], dim=1)
        inv_full = torch.cat([take(clip["inversion"][int(step)]), extra], dim=1)
        x_in = torch.cat([x_full, x_full, inv_full])
        c_in = torch.cat([take(clip["uc"]), take(clip["c"]), take(clip["target_cond"])])
        with torch.no_grad():
            e_u, e_c, _ = port.unet_forward(sd, x_in, torch.cat([tt] * 3), c_in, 8, hooks).chunk(3)
        tb = ok.make_schedule(DDIM_STEPS)
        idx = DDIM_STEPS - 1 - i
        xp, _ = ok.ddim_cfg_step(x.numpy(), e_u.numpy(), e_c.numpy(), tb["ddim_alphas"][idx], tb["ddim_alphas_prev"][idx],
                                 tb["ddim_sigmas"][idx], tb["ddim_sqrt_one_minus_alphas"][idx], CFG_SCALE)
        x = torch.from_numpy(xp)
        ts.append(time.perf_counter() - t0)
    return float(np.mean(ts[warm:])), ts


def run_reference_arm(args, rank):
    if rank != 0:
        return
    from vface_b200 import synth
    from vface_b200.latent_diffusion import LatentDiffusion
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = synth.synth_state_dict(LatentDiffusion().model.diffusion_model.state_dict(), seed=1)
    frames = CPU_ARM_FRAMES
    # bounded: two frames per step (so that the flow warp between frames is exercised); at most ~5 minutes in total
    t_probe, _ = cpu_step_seconds(sd, frames, 0, 1)
    budget = 300.0
    K, W = args.steps, args.warmup
    fit = max(1, int(budget / max(t_probe, 1e-3)) - 1)
    note = None
    if W + K > fit:
        W = min(W, max(0, fit // 4))
        K2 = max(1, fit - W)
        if K2 < K:
            note = f"steps reduced from {K} to {K2} to bound the CPU run to ~{int(budget)} s"
            K = K2
    t_step, _ = cpu_step_seconds(sd, frames, W, K)
    value = frames / (DDIM_STEPS * t_step)
    sample = (f"{frames} frame(s) per step (UNet batch {3 * frames}, hooks on), {K} timed DDIM steps of the same "
              f"full-size UNet; fp32 PyTorch on {cores} host threads")
    line = dict(impl="reference", metric=METRIC, value=value, unit=UNIT, n_gpus=args.gpus, steps=K, warmup=W,
                ms_per_step=t_step * 1e3, higher_is_better=True, scaling="weak", vs_baseline=None, dtype="f32",
                data="synthetic",
                config=workload_config(args.frames, max(args.gpus, 1)),
                cpu_baseline=dict(value=value, unit=UNIT, cores=cores, kind="port", sample=sample),
                e2e=dict(value=value, unit=UNIT, h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    if note:
        line["note"] = note
    print(json.dumps(line), flush=True)


# ---- per-kernel rooflines of the memory-bound kernels (timed live inside steps of the bench loop) ---------------------
def hbm_write_ceiling_gbs(device):
    """Pure-write bandwidth of this GPU, measured live (fill of 768 MiB, best of 5, CUDA events).  The copy peak of
    MEASURED_PEAKS.json is half reads, half writes; a kernel that mostly WRITES (the QKV projection: 1 row-unit in, 3 out)
    is bounded by this lower figure (profiles/r2_write_bw.txt: 3.84 TB/s against 6.56 TB/s of copy traffic)."""
    buf = torch.empty(768 << 20, dtype=torch.uint8, device=device)
    best = 1e9
    for _ in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        buf.zero_()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return buf.numel() / (best * 1e-3) / 1e9


def summarize_secondary(events, n_steps, pk, write_gbs=None):
    """{kind: [(ms, work, unit), ...]} from ops.profile_kernels -> one row per kernel kind.  `work` is the ALGORITHMIC bytes
    (or flops) of the launch as DESIGN.md section 3 defines them; fractions are of the measured peaks."""
    import re
    rows = []
    for kind, evs in sorted(events.items()):
        ms = float(np.sum([e[0] for e in evs]))
        work = float(np.sum([e[1] for e in evs]))
        unit = evs[0][2]
        if ms <= 0:
            continue
        if unit == "B":
            ach = work / (ms * 1e-3) / 1e9
            row = dict(kernel=kind, bound="hbm", launches_per_step=len(evs) / n_steps, ms_per_step=ms / n_steps,
                       algorithmic_bytes_per_step=work / n_steps, achieved=ach, unit="GB/s", peak=pk["hbm"], frac=ach / pk["hbm"])
            m = re.match(r"linear_proj k=(\d+) n=(\d+)", kind)
            if m and write_gbs:
                # write-heavy projection: its floor is max(all bytes at the copy peak, written bytes at the write ceiling)
                k_, n_ = int(m.group(1)), int(m.group(2))
                wfrac = n_ / (k_ + n_ + (n_ if "+res" in kind else 0))
                floor_ms = max(work / pk["hbm"], work * wfrac / write_gbs) / 1e9 * 1e3
                row.update(written_bytes_per_step=work * wfrac / n_steps, hbm_write_ceiling=write_gbs, frac_of_floor=floor_ms / ms)
            rows.append(row)
        else:
            ach = work / (ms * 1e-3) / 1e12
            rows.append(dict(kernel=kind, bound="tensor", launches_per_step=len(evs) / n_steps, ms_per_step=ms / n_steps,
                             algorithmic_flops_per_step=work / n_steps, achieved=ach, unit="TFLOP/s", peak=pk["tc_sustained"],
                             frac=ach / pk["tc_sustained"]))
    rows.sort(key=lambda r: -r["ms_per_step"])
    return dict(peak_source=f"{pk['src']} (MEASURED_PEAKS.json): HBM copy {pk['hbm']} GB/s, bf16 sustained {pk['tc_sustained']} TF/s",
                how="CUDA event pair around every vface_b200 launch during 2 extra steps of the timed loop (same shapes, same stream)",
                kernels=rows)


# ---- sharded == unsharded ----------------------------------------------------------------------
def shard_vs_single(sampler, rank, world, device, frames_per_rank, ddim_steps, seed=23):
    """ONE clip of frames_per_rank * world frames through DDIMSampler.sample twice: sharded by contiguous frame chunk over
    the ranks (halo over NCCL), and whole on every rank (no shard active).  Returns the comparison (rank 0's view)."""
    import torch.distributed as dist
    from vface_b200 import synth
    total = frames_per_rank * world
    sampler.make_schedule(ddim_steps, ddim_eta=0.0, verbose=False)
    clip = synth.synth_clip(total, seed=seed, steps=sampler.ddim_timesteps)
    g = lambda t: t.to(device)

    def sample(sh):
        from vface_b200 import frame_shard
        frame_shard.activate(sh)
        take = (lambda t: t) if sh is None else sh.take
        flow = clip["flow"] if sh is None else sh.local_flow(clip["flow"])
        inv = {t: g(take(v)) for t, v in clip["inversion"].items()}
        out, _ = sampler.sample(
            S=ddim_steps, batch_size=total if sh is None else sh.frames, shape=(4, 64, 64), conditioning=g(take(clip["c"])),
            target_conditioning=g(take(clip["target_cond"])), inverse_results_dir=inv, x_T=g(take(clip["x_T"])),
            flow=[g(f) for f in flow], unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=g(take(clip["uc"])),
            eta=0.0, verbose=False,
            test_model_kwargs=dict(inpaint_image=g(take(clip["inpaint_image"])), inpaint_mask=g(take(clip["inpaint_mask"]))))
        return out

    from vface_b200 import frame_shard
    sh = frame_shard.FrameShard(rank, world, total)
    with torch.no_grad():
        part = sample(sh)
        gathered = sh.gather_frames(part)
        whole = sample(None)
    frame_shard.activate(None)
    sampler.make_schedule(DDIM_STEPS, ddim_eta=0.0, verbose=False)
    diff = (gathered.double() - whole.double())
    rel = (diff.norm() / whole.double().norm()).item()
    # the ranks computed `whole` independently on identical inputs: it must agree bit for bit across GPUs
    ref = whole.clone()
    if world > 1:
        dist.broadcast(ref, 0)
    same_whole = torch.tensor([float(torch.equal(ref, whole))], device=device)
    if world > 1:
        dist.all_reduce(same_whole, op=dist.ReduceOp.MIN)
    own = slice(sh.lo, sh.hi)
    return dict(total_frames=total, frames_per_rank=frames_per_rank, ddim_steps=ddim_steps, dtype=str(whole.dtype),
                model_dtype=str(next(sampler.model.model.diffusion_model.parameters()).dtype),
                rel_l2_sharded_vs_single=rel, max_abs=diff.abs().max().item(), bit_exact=bool(torch.equal(gathered, whole)),
                own_frames_bit_exact=bool(torch.equal(part, whole[own])),
                single_gpu_runs_identical_across_ranks=bool(same_whole.item() == 1.0),
                halo_messages=sh.halo_messages, shard_bounds=frame_shard.shard_bounds(total, world))


def verify_shard(rank, world, device):
    """--verify-shard: the strict form of D1 in the fp32 reference-precision path (and bf16 beside it): a 2-frames-per-rank
    clip, 4 DDIM steps (3 is not a valid DDIM-step count for the reference's schedule) with hooks and flow,
    sharded over the ranks vs whole on one GPU; fp32 must agree to <= 1e-5."""
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    res = {}
    ok = True
    for name, dt in (("fp32", torch.float32), ("bf16", torch.bfloat16)):
        model, _ = build_model(device, dt)
        sampler = DDIMSampler(model)
        r = shard_vs_single(sampler, rank, world, device, frames_per_rank=2, ddim_steps=4)
        res[name] = r
        if name == "fp32":
            ok = ok and r["rel_l2_sharded_vs_single"] <= 1e-5
        del model, sampler
        torch.cuda.empty_cache()
    if rank == 0:
        print(json.dumps(dict(verify_shard=res, n_gpus=world, passed=bool(ok))), flush=True)
    if not ok:
        sys.exit(1)


# ---- main arm ---------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--frames", type=int, default=32, help="frames per GPU")
    ap.add_argument("--impl", default="vface_b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--elide-recon", action="store_true", help="skip the output-dead recon branch (SURVEY.md F3)")
    ap.add_argument("--no-elide-extra", action="store_true", help="skip the secondary elided-recon measurement")
    ap.add_argument("--total-frames", type=int, default=0,
                    help="strong scaling: ONE clip of this many frames split over the ranks (BASELINE.json configs[4]: 256); "
                         "overrides --frames")
    ap.add_argument("--no-clip256", action="store_true", help="skip the secondary 256-frame strong-scaling measurement")
    ap.add_argument("--no-secondary", action="store_true", help="skip the per-kernel roofline_secondary pass")
    ap.add_argument("--no-graph-extra", action="store_true", help="skip the secondary CUDA-graph-replay measurement (N=1 only)")
    ap.add_argument("--verify-shard", action="store_true",
                    help="correctness only (SURVEY.md section 4, D1): one clip sharded over the ranks == the same clip on one rank")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch.distributed as dist
    from vface_b200 import frame_shard, ops
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler

    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: vface_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # NCCL's version banner must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=device)
    W = max(args.warmup, 3)
    K = max(args.steps, 1)
    strong = args.total_frames > 0
    if strong:
        if args.total_frames % world:
            sys.exit(f"--total-frames {args.total_frames} must divide by the {world} ranks")
        frames = args.total_frames // world
    else:
        frames = args.frames
    total_frames = frames * world
    if args.verify_shard:
        verify_shard(rank, world, device)
        if world > 1:
            dist.destroy_process_group()
        return

    model, sd = build_model(device, torch.bfloat16)
    sampler = DDIMSampler(model, elide_dead_recon=args.elide_recon)
    sampler.make_schedule(DDIM_STEPS, ddim_eta=0.0, verbose=False)
    steps = sampler.ddim_timesteps
    time_range = np.flip(steps)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    class Runner:
        """One rank's shard (`n` frames) of a clip of n * world frames, resident in HBM, and the step over it."""

        def __init__(self, n, pinned=True):
            self.n = n
            self.shard = frame_shard.FrameShard(rank, world, n * world)
            clip = local_clip(n, rank, steps)
            pin = (lambda t: t.pin_memory()) if pinned else (lambda t: t)
            self.host = {k: pin(v) for k, v in clip.items() if isinstance(v, torch.Tensor)}
            self.host_flow = pin(torch.cat(clip["flow"], dim=0))
            self.host_inv = {t: pin(v) for t, v in clip["inversion"].items()}
            self.dev = {k: v.to(device) for k, v in self.host.items()}
            self.dev_flow = self.host_flow.to(device)
            self.dev_inv = {t: v.to(device) for t, v in self.host_inv.items()}
            self.kw = dict(test_model_kwargs=dict(inpaint_image=self.dev["inpaint_image"], inpaint_mask=self.dev["inpaint_mask"]))

        def activate(self):
            frame_shard.activate(self.shard)
            sampler._register_hooks(self.dev_flow)
            sampler._inv_cache = self.dev_inv

        def step(self, i, x):
            i = i % DDIM_STEPS
            step = int(time_range[i])
            ts = torch.full((self.n,), step, device=device, dtype=torch.long)
            x_prev, _ = sampler.p_sample_ddim_with_inverse(
                x, self.dev["c"], ts, index=DDIM_STEPS - 1 - i, target_conditioning=self.dev["target_cond"],
                inverse_results_dir=self.dev_inv, unconditional_guidance_scale=CFG_SCALE,
                unconditional_conditioning=self.dev["uc"], flow=self.dev_flow, _step=step, **self.kw)
            return x_prev

        def timed(self, warm, k, first=0):
            """`warm` untimed then `k` timed steps; returns (ms of the k steps on this rank, last latents)."""
            x = self.dev["x_T"]
            with torch.no_grad():
                for i in range(first, first + warm):
                    x = self.step(i, x)
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(first + warm, first + warm + k):
                    x = self.step(i, x)
                e1.record()
                barrier()
            return e0.elapsed_time(e1), x

    def max_over_ranks(v):
        t = torch.tensor([float(v)], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item()

    run = Runner(frames)
    run.activate()
    shard = run.shard
    host, host_flow, host_inv = run.host, run.host_flow, run.host_inv
    one_step = run.step

    x = run.dev["x_T"]
    with torch.no_grad():
        for i in range(W):
            x = one_step(i, x)
        barrier()
        clocks = ClockSampler(physical_gpu_index(local_rank))
        clocks.start()
        ops.profile_attention(min_tokens=4096)
        launches0 = ops.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for i in range(W, W + K):
            x = one_step(i, x)
        ev1.record()
        barrier()
        clock_info = clocks.stop()
        launches = ops.launch_count - launches0
        attn_ms = ops.profile_attention(None)
    ms_total = ev0.elapsed_time(ev1)

    # ---- memory-bound kernels, timed live inside two more steps of the same loop (not inside the K timed steps: an event
    # pair around each of ~230 launches per step would perturb the headline) --------------------------------------------
    secondary = None
    if not args.no_secondary:
        ops.profile_kernels(True)
        with torch.no_grad():
            for i in range(W + K, W + K + 2):
                x = one_step(i, x)
        secondary = summarize_secondary(ops.profile_kernels(None), 2, peaks(), hbm_write_ceiling_gbs(device))
        barrier()

    # secondary figure ("report throughput both ways"): the same steps with the output-dead recon
    # branch skipped (UNet batch 2 x frames).  Never the headline: `value` above is the faithful 3-branch step.
    elide_ms = None
    if not args.elide_recon and not args.no_elide_extra:
        sampler.elide_dead_recon = True
        run.activate()
        ke = min(K, 5)
        t_e, _ = run.timed(2, ke)
        elide_ms = t_e / ke
        sampler.elide_dead_recon = False
        run.activate()
    # secondary figure: the same steps replayed as CUDA graphs (one graph per schedule position, captured on
    # first use by DDIMSampler(cuda_graphs=True)).  Five schedule positions are captured, then replayed and timed.
    graph_ms = None
    if world == 1 and not args.no_graph_extra:
        try:
            kg = 5
            common = dict(use_original_steps=False, quantize_denoised=False, temperature=1., noise_dropout=0., score_corrector=None,
                          corrector_kwargs=None, unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=run.dev["uc"], **run.kw)

            def graph_steps(xg):
                for i in range(kg):
                    xg, _ = sampler._graphed_step(xg, run.dev["c"], int(time_range[i]), DDIM_STEPS - 1 - i, run.dev["target_cond"],
                                                  run.dev_flow, common)
                return xg
            with torch.no_grad():
                xg = graph_steps(run.dev["x_T"])             # capture (with an eager warm-up each)
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                xg = graph_steps(run.dev["x_T"])             # replay
                g1.record()
                barrier()
                xe = run.dev["x_T"]
                for i in range(kg):
                    xe = one_step(i, xe)
                barrier()
            graph_ms = dict(ms_per_step=g0.elapsed_time(g1) / kg, steps=kg, value=total_frames / (DDIM_STEPS * g0.elapsed_time(g1) / kg * 1e-3),
                            unit=UNIT, identical_to_eager=bool(torch.equal(xg, xe)),
                            note="the first 5 schedule positions replayed as CUDA graphs (DDIMSampler(cuda_graphs=True)); secondary figure")
            sampler._graphs.clear()
            torch.cuda.empty_cache()
        except Exception as e:            # a secondary figure must never take the bench line down
            graph_ms = dict(error=repr(e)[:300])
    lsum = torch.tensor([float(launches)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(lsum, op=dist.ReduceOp.SUM)
    ms_per_step = max_over_ranks(ms_total) / K
    value = total_frames / (DDIM_STEPS * ms_per_step * 1e-3)
    elide = None
    if elide_ms is not None:
        te = max_over_ranks(elide_ms)
        elide = dict(value=total_frames / (DDIM_STEPS * te * 1e-3), unit=UNIT, ms_per_step=te, branches=2,
                     note="same step with the recon branch skipped: bit-identical samples (SURVEY.md F3), secondary figure only")

    # ---- BASELINE.json configs[4] as written: ONE 256-frame clip split over the ranks (128 / 64 / 32 frames per rank at
    # 2 / 4 / 8 GPUs; all 256 on one GPU at N=1), so that the driver's N = 1, 2, 4, 8 lines carry a STRONG-scaling series
    # next to the weak-scaling headline.  Secondary figure; same step, same kernels. -------------------------------------
    clip256 = None
    if not strong and not args.no_clip256 and 256 % world == 0:
        n256 = 256 // world
        if n256 == frames:
            clip256 = dict(total_frames=256, frames_per_gpu=n256, ms_per_step=ms_per_step, value=value, unit=UNIT,
                           scaling="strong", note="identical to the headline run at this N")
        else:
            del x
            torch.cuda.empty_cache()
            r256 = Runner(n256, pinned=False)
            r256.activate()
            k256 = 3
            t256, x256 = r256.timed(2, k256)
            ms256 = max_over_ranks(t256) / k256
            clip256 = dict(total_frames=256, frames_per_gpu=n256, ms_per_step=ms256, steps=k256, warmup=2,
                           value=256 / (DDIM_STEPS * ms256 * 1e-3), unit=UNIT, scaling="strong",
                           halo=dict(messages=r256.shard.halo_messages, bytes=r256.shard.halo_bytes) if world > 1 else None,
                           note="one 256-frame clip split by contiguous frame chunk over the ranks; efficiency(N) = value(N) / (N * value(1)) "
                                "with value(1) = this object in the N=1 line")
            del r256, x256
            torch.cuda.empty_cache()
            run.activate()

    # ---- sharded == unsharded on real ranks: one small clip through DDIMSampler.sample, sharded
    # over the ranks (halo over NCCL) vs whole on every rank.  fp32 reference-precision path (a second, fp32 copy of the
    # UNet for the duration of the check): the two evaluations are the same computation up to the batch-size-dependent
    # kernel choices of cuDNN / cuBLAS, i.e. fp32 round-off; in bf16 they would be two independent roundings ~1e-2 apart,
    # which proves nothing.  --verify-shard runs a longer form and asserts. -----------------------------------------------
    shard_check = None
    if world > 1:
        m32, _ = build_model(device, torch.float32, state=sd)
        shard_check = shard_vs_single(DDIMSampler(m32), rank, world, device, frames_per_rank=2, ddim_steps=2)
        shard_check["passed"] = bool(shard_check["rel_l2_sharded_vs_single"] <= 1e-5)
        del m32
        torch.cuda.empty_cache()
        run.activate()

    # ---- roofline of the dominant kernel (fused attention, N=4096) -------------------------------------
    pk = peaks()
    n_branches = 2 if args.elide_recon else 3
    roof = None
    if attn_ms:
        avg_ms = float(np.mean(attn_ms))
        flops = ATTN4096_FLOPS_PER_SAMPLE * frames * n_branches
        ach = flops / (avg_ms * 1e-3) / 1e12
        traffic = None
        tp = os.path.join(ROOT, "profiles", "attn_traffic.json")
        if os.path.exists(tp):
            try:
                traffic = json.load(open(tp)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        # what the hardware allows at d = 40: one MUFU.EX2 per 160 MMA flops, 15.9 MUFU.EX2/clk/SM measured
        # (experiments/mufu_rate.cu), at the SM clock sampled during the timed steps
        sm_mhz = (clock_info or {}).get("sm_mhz") or pk.get("sm_max_mhz") or 1965.0
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count
        mufu_ceiling = n_sm * 15.9 * sm_mhz * 1e6 * 160.0 / 1e12
        roof = dict(bound="tensor", kernel="vf::attn_tc_kernel<BN=64, S/O 128 + P 32 TMEM cols, 3 CTAs/SM, 2 MMA issuer warps> (N=4096, 8 heads x d40)", achieved=ach,
                    peak=pk["tc_sustained"], unit="TFLOP/s", frac=ach / pk["tc_sustained"], traffic=traffic,
                    peak_source=f"{pk['src']} sustained bf16 (kernel timed inside a long step); burst {pk['tc_burst']}",
                    frac_of_burst=ach / pk["tc_burst"], launches_timed=len(attn_ms), avg_launch_ms=avg_ms,
                    mufu_ex2_ceiling_tflops=mufu_ceiling, frac_of_mufu_ex2_ceiling=ach / mufu_ceiling,
                    ceiling_note="d_head 40: one exponential per 160 MMA flops; ceiling = SMs x 15.9 MUFU.EX2/clk x sampled SM clock x 160",
                    share_of_step=float(np.sum(attn_ms)) / K / ms_per_step,
                    algorithmic_flops_per_launch=flops)

    # ---- end to end through the public API, host buffers ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        h2d = sum(v.numel() * v.element_size() for v in host.values()) + host_flow.numel() * 4 \
            + sum(v.numel() * 4 for v in host_inv.values())
        d2h = frames * 4 * 64 * 64 * 4
        out_host = torch.empty(frames, 4, 64, 64).pin_memory()

        def sample_e2e():
            g = lambda t: t.to(device, non_blocking=True)
            inv = {t: g(v) for t, v in host_inv.items()}
            samples, _ = sampler.sample(
                S=DDIM_STEPS, batch_size=frames, shape=(4, 64, 64), conditioning=g(host["c"]),
                target_conditioning=g(host["target_cond"]), inverse_results_dir=inv, x_T=g(host["x_T"]),
                flow=g(host_flow), unconditional_guidance_scale=CFG_SCALE, unconditional_conditioning=g(host["uc"]),
                eta=0.0, verbose=False,
                test_model_kwargs=dict(inpaint_image=g(host["inpaint_image"]), inpaint_mask=g(host["inpaint_mask"])))
            out_host.copy_(samples, non_blocking=True)
            torch.cuda.synchronize()
            return out_host

        barrier()
        t0 = time.perf_counter()
        sample_e2e()
        barrier()
        t_e2e = time.perf_counter() - t0
        te = torch.tensor([t_e2e], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e = dict(value=total_frames / te.item(), unit=UNIT, h2d_bytes_per_step=int(h2d // DDIM_STEPS),
                   d2h_bytes_per_step=int(d2h // DDIM_STEPS), seconds=te.item(),
                   call="DDIMSampler.sample(S=50, ...) on pinned host buffers, one full DDIM-50 pass")

    # ---- CPU baseline beside it (rank 0, N=1 only) -----------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        sd32 = {k: v.float() for k, v in sd.items()}
        cpu_frames = CPU_ARM_FRAMES        # two frames: the flow warp between them is part of the sample
        t_step, _ = cpu_step_seconds(sd32, cpu_frames, 1, 1)
        cpu = dict(value=cpu_frames / (DDIM_STEPS * t_step), unit=UNIT, cores=cores, kind="port",
                   sample=f"{cpu_frames} frames x 1 timed DDIM step after 1 warm-up step (UNet batch {3 * cpu_frames}, hooks on, "
                          f"flow warp between the frames) of the same full-size UNet, fp32 PyTorch on {cores} host threads; "
                          f"{t_step:.1f} s per step")

    if rank == 0:
        line = dict(metric=METRIC, value=value, unit=UNIT, n_gpus=world, steps=K, warmup=W, ms_per_step=ms_per_step,
                    higher_is_better=True, scaling="strong" if strong else "weak", vs_baseline=None, dtype="bf16",
                    data="synthetic", config=workload_config(frames, world, n_branches, strong),
                    clocks=clock_info, e2e=e2e, gpu_launches=int(lsum.item()), roofline=roof,
                    roofline_secondary=secondary, cpu_baseline=cpu, elide_dead_recon=elide, clip256=clip256, cuda_graph=graph_ms,
                    shard_check=shard_check,
                    halo=dict(messages=shard.halo_messages, bytes=shard.halo_bytes) if world > 1 else None)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
