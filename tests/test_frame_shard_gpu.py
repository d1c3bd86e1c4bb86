"""D1: frame-sharded sampling == unsharded sampling.

The driver's GPU box may expose a single device, so the two ranks are emulated one after the other
on cuda:0 with a FrameShard whose halo exchange replays what rank 0 would have sent over NCCL (the
real NCCL path is exercised by `bench.py --gpus 2`; the gloo path by tests/test_host_logic.py).
Shard assignment is bit-exact by construction (shard_bounds); latents agree to fp32 round-off
(cuBLAS/cuDNN pick batch-size-dependent kernels, so bitwise equality across batch sizes is not defined).
In bf16 the two runs are two independent roundings of the same fp32 trajectory (each within the 1e-2
per-step budget of the fp32 result on this random-weight UNet with S=4), so they are compared at 2e-2;
the fp32 case (1e-5) is the one that proves the sharded evaluation is the same computation.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

SMALL = dict(model_channels=32, num_heads=2)


class ReplayShard:
    """FrameShard stand-in: rank 0 records its halo sends, rank 1 replays them in order."""

    def __init__(self, rank, world, total, tape):
        from vface_b200.frame_shard import FrameShard
        self._fs = FrameShard(rank, world, total)
        self.rank, self.world_size, self.total_frames = rank, world, total
        self.lo, self.hi = self._fs.lo, self._fs.hi
        self.tape = tape
        self.pos = 0

    def take(self, t):
        return self._fs.take(t)

    def local_flow(self, flow):
        return self._fs.local_flow(flow)

    def exchange_halo_async(self, q_last, k_last):
        from vface_b200.frame_shard import _PendingHalo
        return _PendingHalo(None, self.exchange_halo(q_last, k_last), None)

    def exchange_halo(self, q_last, k_last):
        if self.rank == 0:
            self.tape.append((q_last.clone(), k_last.clone()))
            return None, None
        q, k = self.tape[self.pos]
        self.pos += 1
        return q, k


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-5), (torch.bfloat16, 2e-2)])
def test_two_shards_equal_unsharded(dtype, tol):
    from oracle import kernels as ok
    from tests.test_pipeline_gpu import build, rel_l2, run_sample
    from vface_b200 import frame_shard, synth
    S, F = 4, 4
    steps = ok.make_schedule(S)["ddim_timesteps"]
    clip = synth.synth_clip(F, steps=steps)
    _, sampler, _ = build(SMALL, dtype)
    frame_shard.activate(None)
    full, _ = run_sample(sampler, clip, S, F, clip["inversion"])

    tape, parts = [], []
    try:
        for rank in range(2):
            sh = ReplayShard(rank, 2, F, tape)
            frame_shard.activate(sh)
            local = {k: (sh.take(v) if isinstance(v, torch.Tensor) else v) for k, v in clip.items()}
            local["flow"] = sh.local_flow(clip["flow"])
            inv = {t: sh.take(v) for t, v in clip["inversion"].items()}
            out, _ = run_sample(sampler, local, S, sh.hi - sh.lo, inv, flow=local["flow"])
            parts.append(out)
    finally:
        frame_shard.activate(None)
    # 2 flow-active modules x (q,k packed in one message) x S steps
    assert len(tape) == 2 * S
    assert rel_l2(torch.cat(parts), full) < tol


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL ranks)")
def test_sharded_equals_unsharded_on_real_ranks():
    """SURVEY.md section 4, D1 on hardware: ONE clip sharded by frame chunk over real NCCL ranks (one process per GPU, halo
    over NVLink) reproduces the same clip run whole on one GPU -- fp32 path to <= 1e-5 relative L2.  Runs
    `bench.py --verify-shard` under torchrun on min(device_count, 8) GPUs; skipped on a one-GPU box."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    n = 1 << (min(torch.cuda.device_count(), 8).bit_length() - 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", "29617", os.path.join(root, "bench.py"), "--gpus", str(n), "--verify-shard"]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, cwd=root)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    line = [ln for ln in res.stdout.splitlines() if ln.startswith("{") and "verify_shard" in ln][-1]
    out = json.loads(line)
    assert out["passed"] and out["n_gpus"] == n
    fp32 = out["verify_shard"]["fp32"]
    assert fp32["rel_l2_sharded_vs_single"] <= 1e-5 and fp32["halo_messages"] > 0
    assert fp32["single_gpu_runs_identical_across_ranks"]
    assert fp32["shard_bounds"] == [[2 * r, 2 * r + 2] for r in range(n)]
