"""Host-side logic that needs neither a GPU nor the reference: the C-ABI library's exported symbols,
frame sharding + halo exchange over gloo (world_size 2), the sampler's schedule tables, the drop-in
installer and the loud failure without CUDA."""
import ctypes
import os
import re
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- C-ABI ------------------------------------------------------------------------------------------
def _header_symbols():
    src = open(os.path.join(ROOT, "include", "vface_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(vf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from vface_b200 import _lib, build
    build.build()                                  # no-op when fresh; nvcc cross-compiles without a GPU
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _header_symbols()
    assert len(syms) >= 8
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/vface_b200.h but not exported"
    assert set(syms) == set(_lib.SIGNATURES), "ctypes signature table out of sync with the header"
    lib.vf_abi_version.restype = ctypes.c_int
    assert lib.vf_abi_version() == _lib.ABI_VERSION


def test_ops_refuse_cpu_tensors():
    from vface_b200 import ops
    x = torch.zeros(1, 8, 64)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.attention(x, x, x, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.fsai_blend(x, x.clone())
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.flow_warp_blend(torch.zeros(2, 64, 8), torch.zeros(1, 2, 8, 8), 0.8, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.ddim_cfg_step(torch.zeros(4), torch.zeros(4), torch.zeros(4), 0.5, 0.6, 0.0, 0.7, 3.0)


def test_product_does_not_import_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|^\s*from\s+\.+\s*oracle\b", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "vface_b200")):
        for f in files:
            if f.endswith(".py"):
                assert not pat.search(open(os.path.join(base, f)).read()), f


# ---- schedule tables ----------------------------------------------------------------------------------
@pytest.mark.parametrize("S,eta", [(10, 0.0), (10, 0.5), (50, 0.0), (50, 0.5)])
def test_sampler_tables_vs_reference_golden(S, eta):
    from vface_b200.latent_diffusion import LatentDiffusion
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    g = np.load(os.path.join(GOLD, "schedule.npz"))
    s = DDIMSampler(LatentDiffusion(unet=torch.nn.Identity()))
    s.make_schedule(S, ddim_eta=eta, verbose=False)
    tag = f"S{S}_eta{int(eta * 10)}"
    assert np.array_equal(s.ddim_timesteps, g[f"timesteps_{tag}"])
    for k, hk in (("ddim_alphas", "a_t"), ("ddim_alphas_prev", "a_prev"), ("ddim_sigmas", "sigma"),
                  ("ddim_sqrt_one_minus_alphas", "s1m")):
        assert np.array_equal(s._host_tables[hk], g[f"{k}_{tag}"]), k
    assert np.array_equal(s._host_tables["acp"], g["alphas_cumprod"])


def test_bad_step_count_fails_like_reference():
    """S=3 -> range(0,1000,333)+1 reaches 1000 -> IndexError in the reference."""
    from vface_b200.latent_diffusion import LatentDiffusion
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    s = DDIMSampler(LatentDiffusion(unet=torch.nn.Identity()))
    with pytest.raises(IndexError):
        s.make_schedule(3, verbose=False)


# ---- hooks registration (plugin mechanism) --------------------------------------------------------------
def test_hook_registration_census():
    """input 6 / middle 1 / output 9 attn1 modules; instance-level forward patching."""
    from vface_b200.latent_diffusion import LatentDiffusion
    from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
    from vface_b200.ldm.models.pnp_utils import find_all_modules_by_name, register_spa_attn_injection
    model = LatentDiffusion(unet_config=dict(model_channels=32, num_heads=2))
    unet = model.model.diffusion_model
    census = [len(find_all_modules_by_name(b, "attn1")[0]) for b in (unet.input_blocks, unet.middle_block, unet.output_blocks)]
    assert census == [6, 1, 9]
    names = find_all_modules_by_name(unet.input_blocks, "attn1")[1]
    assert names == [f"{i}.1.transformer_blocks.0.attn1" for i in (1, 2, 4, 5, 7, 8)]
    s = DDIMSampler(model)
    s._register_hooks(flow=None)
    for blocks, n in ((unet.input_blocks, 6), (unet.middle_block, 1), (unet.output_blocks, 9)):
        mods = find_all_modules_by_name(blocks, "attn1")[0]
        assert sum("forward" in m.__dict__ for m in mods) == n
    with pytest.raises(NotImplementedError):
        register_spa_attn_injection(s, 1, fusion="adaIn")
    # block_indices filter: the others only get the schedule attribute, like the reference
    model2 = LatentDiffusion(unet_config=dict(model_channels=32, num_heads=2))
    s2 = DDIMSampler(model2)
    register_spa_attn_injection(s2, 7, input_blocks=True, output_blocks=False, block_indices=[0, 2], fusion="fft")
    mods = find_all_modules_by_name(model2.model.diffusion_model.input_blocks, "attn1")[0]
    assert ["forward" in m.__dict__ for m in mods] == [True, False, True, False, False, False]
    assert mods[1].injection_schedule == 7


def test_install_rebinds_reference_modules():
    import types
    import vface_b200
    fake = {}
    for name in ("ldm", "ldm.models", "ldm.models.diffusion", "ldm.models.diffusion.ddim_w_inv", "ldm.models.pnp_utils"):
        if name in sys.modules:
            pytest.skip("a real ldm package is already imported in this process")
    try:
        for name in ("ldm", "ldm.models", "ldm.models.diffusion", "ldm.models.diffusion.ddim_w_inv", "ldm.models.pnp_utils"):
            fake[name] = sys.modules[name] = types.ModuleType(name)
        done = vface_b200.install()
        from vface_b200.ldm.models.diffusion.ddim_w_inv import DDIMSampler
        assert fake["ldm.models.diffusion.ddim_w_inv"].DDIMSampler is DDIMSampler
        assert ("ldm.models.pnp_utils", "register_spa_attn_injection") in done
    finally:
        for name in fake:
            sys.modules.pop(name, None)


# ---- frame sharding ---------------------------------------------------------------------------------
def test_shard_bounds_bit_exact():
    from vface_b200.frame_shard import shard_bounds
    assert shard_bounds(256, 8) == [(32 * r, 32 * (r + 1)) for r in range(8)]
    assert shard_bounds(256, 2) == [(0, 128), (128, 256)]
    assert shard_bounds(10, 4) == [(0, 2), (2, 5), (5, 7), (7, 10)]
    for f, g in ((256, 8), (33, 4), (7, 7)):
        b = shard_bounds(f, g)
        assert b[0][0] == 0 and b[-1][1] == f and all(b[i][1] == b[i + 1][0] for i in range(g - 1))
    with pytest.raises(ValueError):
        shard_bounds(3, 4)


def _halo_worker(rank, world, port, frames_total, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kernels as ok
        from vface_b200.frame_shard import FrameShard
        h = w = 16
        c = 8
        g = torch.Generator().manual_seed(0)
        x = torch.randn(frames_total, h * w, c, generator=g)
        flow = torch.randn(frames_total - 1, 2, h, w, generator=g) * 2
        full = ok.flow_warp_blend(x.numpy(), flow.numpy(), 0.8, h, w)
        sh = FrameShard(rank, world, frames_total)
        xl = sh.take(x)
        fl = sh.local_flow(flow)
        halo_q, halo_k = sh.exchange_halo_async(xl[-1], -xl[-1]).wait()     # synchronous on CPU tensors
        if rank == 0:
            assert halo_q is None and len(fl) == sh.frames - 1
            loc = ok.flow_warp_blend(xl.numpy(), fl.numpy(), 0.8, h, w)
        else:
            assert torch.equal(halo_q, x[sh.lo - 1]) and torch.equal(halo_k, -x[sh.lo - 1])
            assert len(fl) == sh.frames
            loc = ok.flow_warp_blend(xl.numpy(), fl.numpy(), 0.8, h, w, prev_halo=halo_q.numpy())
        assert np.array_equal(loc, full[sh.lo:sh.hi])                 # sharded == unsharded, bit-exact
        gathered = sh.gather_frames(torch.from_numpy(loc))
        assert np.array_equal(gathered.numpy(), full)
        assert sh.halo_messages == (1 if rank + 1 < world else 0)
        ret[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,frames", [(2, 6), (2, 7), (3, 8)])
def test_halo_exchange_gloo(world, frames):
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_halo_worker, args=(world, port, frames, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world))


def test_vae_decoder_state_dict_keys_match_reference_golden():
    """The first-stage decoder mirror has exactly the reference's parameter names (so `first_stage_model.*`
    of the reference checkpoint loads as is); names recorded from the reference in tests/golden/vae_decoder.npz."""
    import os
    import numpy as np
    from vface_b200.ldm.modules.diffusionmodules.model import AutoencoderKLDecoder
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "vae_decoder.npz"))
    small = dict(double_z=True, z_channels=4, resolution=64, in_channels=3, out_ch=3, ch=32, ch_mult=[1, 2, 4, 4],
                 num_res_blocks=2, attn_resolutions=[], dropout=0.0)
    assert sorted(AutoencoderKLDecoder(small).state_dict().keys()) == [str(k) for k in gold["keys"]]
    full = AutoencoderKLDecoder().state_dict()
    assert full["decoder.conv_in.weight"].shape == (512, 4, 3, 3) and full["decoder.mid.attn_1.q.weight"].shape == (512, 512, 1, 1)
    assert full["decoder.conv_out.weight"].shape == (3, 128, 3, 3) and len(full) == 140


def test_flow_adapter_contract():
    """image-resolution flow -> feature-pixel units at feature resolution."""
    import torch
    from vface_b200.scripts import temporal_flow as tf
    # a constant translation of (+16, -8) image pixels at 512x512 is (+2, -1) feature pixels at 64x64
    flow = torch.zeros(3, 2, 512, 512)
    flow[:, 0], flow[:, 1] = 16.0, -8.0
    out = tf.flow_to_feature_resolution(flow, 64)
    assert tuple(out.shape) == (3, 2, 64, 64)
    assert torch.allclose(out[:, 0], torch.full((3, 64, 64), 2.0)) and torch.allclose(out[:, 1], torch.full((3, 64, 64), -1.0))
    # non-square: each component scales with its own axis; the reference's list-of-(1,2,H,W) form is accepted
    lst = [torch.ones(1, 2, 256, 512) * 8.0 for _ in range(2)]
    out = tf.flow_to_feature_resolution(lst, (64, 64))
    assert torch.allclose(out[:, 0], torch.full((2, 64, 64), 1.0)) and torch.allclose(out[:, 1], torch.full((2, 64, 64), 2.0))
    # already at feature resolution: unchanged
    f64 = torch.randn(2, 2, 64, 64)
    assert torch.equal(tf.flow_to_feature_resolution(f64, 64), f64)
    # the authors' commented-out route: resize the frames, estimate at feature resolution
    video = torch.randn(4, 3, 512, 512)
    small = tf.resize_for_flow(video, 8)
    assert tuple(small.shape) == (4, 3, 64, 64)
    assert torch.equal(small, torch.nn.functional.interpolate(video, size=(64, 64), mode="bilinear", align_corners=False))


def test_return_flow_batches_pairs_in_reference_order():
    import torch
    from vface_b200.scripts import temporal_flow as tf
    calls = []

    def estimator(img1, img2, num_flow_updates=20):
        calls.append((img1.clone(), img2.clone(), num_flow_updates))
        # a fake "flow": per-pair mean difference, broadcast; returned as RAFT's list of refinements
        d = (img1 - img2).mean(dim=(1, 2, 3)).view(-1, 1, 1, 1).expand(-1, 2, img1.shape[-2], img1.shape[-1])
        return [d * 0.5, d]

    video = torch.randn(5, 3, 32, 32)
    flow = tf.return_flow(video, estimator)
    assert len(calls) == 1 and tuple(flow.shape) == (4, 2, 32, 32)
    img1, img2, it = calls[0]
    assert it == 20 and torch.equal(img1, video[1:]) and torch.equal(img2, video[:-1])     # compute_flow(frame2, frame1)
    want = (video[1:] - video[:-1]).mean(dim=(1, 2, 3))
    assert torch.allclose(flow[:, 0, 0, 0], want)
    assert tuple(tf.return_flow(video, estimator, feature_size=8).shape) == (4, 2, 8, 8)
    assert tf.return_flow(video[:1], estimator).shape[0] == 0


def test_library_has_no_link_time_cublas_dependency():
    """libvface_b200.so must bind cuBLASLt at run time (dlopen): a DT_NEEDED entry lets the system libcublasLt shadow
    PyTorch's copy when the library is loaded before torch, after which torch's own cublasGemmEx fails."""
    import subprocess
    from vface_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        pytest.skip("library not built")
    out = subprocess.run(["readelf", "-d", _lib.LIB_PATH], capture_output=True, text=True).stdout
    needed = [ln for ln in out.splitlines() if "NEEDED" in ln]
    assert needed, out
    assert not any("cublas" in ln.lower() for ln in needed), needed


def test_projection_fold_algebra_matches_layernorm_then_linear():
    """ops._ProjFold (host side of vf_linear_proj): LN(x) W^T + b == rstd (x Wg^T - mean colsum(Wg)) + (beta W^T + b) with
    Wg = bf16(W o gamma) and the column sums taken over the ROUNDED Wg -- evaluated here in fp64 on the CPU, the same
    expression the kernel's epilogue applies (reference: ldm/modules/attention.py:239 with :172-174)."""
    import torch
    from vface_b200 import ops
    g = torch.Generator().manual_seed(3)
    k, n, rows = 320, 960, 64
    ln = torch.nn.LayerNorm(k).to(torch.bfloat16)
    with torch.no_grad():
        ln.weight.copy_((1.0 + 0.3 * torch.randn(k, generator=g)).to(torch.bfloat16))
        ln.bias.copy_((0.2 * torch.randn(k, generator=g)).to(torch.bfloat16))
    w = (torch.randn(n, k, generator=g) * k ** -0.5).to(torch.bfloat16)
    b = (0.5 * torch.randn(n, generator=g)).to(torch.bfloat16)
    x = (1.5 * torch.randn(rows, k, generator=g) + 0.7).to(torch.bfloat16)
    wg, colsum, b32 = ops._ProjFold.get(w, b, ln)
    assert wg.dtype == torch.bfloat16 and colsum.dtype == torch.float32 and b32.dtype == torch.float32
    assert torch.equal(colsum, wg.float().sum(1))
    xd = x.double()
    mean = xd.mean(1, keepdim=True)
    rstd = 1.0 / torch.sqrt(xd.var(1, unbiased=False, keepdim=True) + ln.eps)
    folded = rstd * (xd @ wg.double().t() - mean * colsum.double()[None]) + b32.double()[None]
    direct = ((xd - mean) * rstd * ln.weight.double() + ln.bias.double()) @ w.double().t() + b.double()
    # the only difference is the bf16 rounding of W o gamma (2^-9 relative per weight, averaged over k = 320 terms)
    assert ((folded - direct).norm() / direct.norm()).item() < 2.0 ** -9
    # cached, and the cache entry pins its source tensors
    assert ops._ProjFold.get(w, b, ln)[0] is wg
    # without a LayerNorm the weight is passed through untouched and the bias only changes dtype
    w2, cs2, b2 = ops._ProjFold.get(w, b, None)
    assert cs2 is None and w2.data_ptr() == w.data_ptr() and torch.equal(b2, b.float())


def test_arming_norm1_does_not_register_a_submodule():
    """BasicTransformerBlock hands norm1 to attn1.project_qkv through a plain attribute: the module tree (state-dict keys,
    which must stay the reference's) is unchanged while it is armed."""
    from vface_b200.ldm.modules.attention import BasicTransformerBlock
    blk = BasicTransformerBlock(320, 8, 40, context_dim=768)
    keys = list(blk.state_dict().keys())
    object.__setattr__(blk.attn1, "_pre_ln", blk.norm1)
    assert list(blk.state_dict().keys()) == keys and "_pre_ln" not in dict(blk.attn1.named_modules())
    object.__setattr__(blk.attn1, "_pre_ln", None)


def test_bench_secondary_roofline_rows_and_write_floor():
    """bench.summarize_secondary: bytes rows against the copy peak, flop rows against the sustained tensor peak, and for the
    projection kernel the second fraction against max(all bytes at the copy peak, written bytes at the WRITE ceiling)."""
    import bench
    pk = dict(hbm=6500.0, tc=1600.0, tc_sustained=1400.0, src="test")
    ev = {
        "linear_proj k=320 n=960+ln": [(0.25, 1.0e9, "B"), (0.25, 1.0e9, "B")],        # 1 GB in 0.25 ms, 3/4 of it written
        "linear_proj k=320 n=320+res": [(0.15, 0.75e9, "B")],
        "layer_norm c=320": [(0.1, 0.5e9, "B")],
        "linear_geglu k=320 n=1280": [(0.5, 0.64e12, "flop")],
    }
    out = bench.summarize_secondary(ev, 2, pk, write_gbs=3840.0)
    rows = {r["kernel"]: r for r in out["kernels"]}
    q = rows["linear_proj k=320 n=960+ln"]
    assert abs(q["achieved"] - 4000.0) < 1e-6 and abs(q["frac"] - 4000.0 / 6500.0) < 1e-9 and q["launches_per_step"] == 1.0
    floor_ms = max(2.0e9 / 6500.0, 2.0e9 * 0.75 / 3840.0) / 1e9 * 1e3          # the write term wins: 0.39 ms for both launches
    assert abs(q["frac_of_floor"] - floor_ms / 0.5) < 1e-9 and q["frac_of_floor"] > q["frac"]
    r = rows["linear_proj k=320 n=320+res"]                                     # 1/3 written: the copy-peak term wins
    assert abs(r["frac_of_floor"] - r["frac"]) < 1e-9
    assert "frac_of_floor" not in rows["layer_norm c=320"]
    g = rows["linear_geglu k=320 n=1280"]
    assert g["bound"] == "tensor" and abs(g["achieved"] - 1280.0) < 1e-6 and abs(g["frac"] - 1280.0 / 1400.0) < 1e-9
