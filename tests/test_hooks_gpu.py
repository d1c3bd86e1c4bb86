"""The plugin mechanism on the GPU: register_spa_attn_injection's patched attn1.forward (pnp_utils.py:92-288)
against the outputs of the UNMODIFIED reference closure (tests/golden/attn_hooks.npz, oracle/make_golden.py::
golden_attn_hooks) in the four reachable configurations: switch off, replace, fft, flow_fix (FSAI + flow warp).
Same seeded weights and inputs as the reference run; fp32 path to 2e-4, bf16 path to the 2e-2 kernel bound.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(dtype):
    from tests.test_oracle_vs_golden import _attn_inputs
    from vface_b200.ldm.modules.attention import CrossAttention
    sd, x = _attn_inputs()
    attn = CrossAttention(query_dim=80, heads=2, dim_head=40)
    attn.load_state_dict({k[2:]: v for k, v in sd.items()})
    attn = attn.cuda().to(dtype).eval()

    class Holder:
        pass
    blk = torch.nn.Module()
    blk.attn1 = attn
    unet = torch.nn.Module()
    unet.input_blocks = torch.nn.ModuleList([blk])
    unet.middle_block = torch.nn.ModuleList([])
    unet.output_blocks = torch.nn.ModuleList([])
    h = Holder(); h.model = Holder(); h.model.model = Holder(); h.model.model.diffusion_model = unet
    return h, attn, x


@pytest.mark.parametrize("mode", ["off", "replace", "fft", "flow_fix"])
@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-4), (torch.bfloat16, 2e-2)])
def test_patched_attn1_closure_vs_reference_golden(mode, dtype, tol):
    from vface_b200.ldm.models.pnp_utils import register_spa_attn_injection
    g = np.load(os.path.join(GOLD, "attn_hooks.npz"))
    h, attn, x = _build(dtype)
    assert np.array_equal(x[:, ::32].numpy(), g["x_rows"])                     # same seeded input as the reference run
    flows = [torch.from_numpy(f)[None] for f in g["flow"]]
    kw = dict(off=dict(switch_on=False, fusion="flow_fix"), replace=dict(switch_on=True, fusion="replace"),
              fft=dict(switch_on=True, fusion="fft", split_ratio_fft=0.8),
              flow_fix=dict(switch_on=True, fusion="flow_fix", split_ratio_fft=0.8, alpha=0.8, flow=flows))[mode]
    register_spa_attn_injection(h, 1, input_blocks=True, output_blocks=False, middle_block=False,
                                attn_component="attn1", chunks=3, **kw)
    assert "forward" in attn.__dict__                                          # instance-level patch, as in the reference
    with torch.no_grad():
        y = attn.forward(x.cuda().to(dtype))
    err = np.abs(y[:, ::32].float().cpu().numpy() - g[f"out_{mode}"]).max()
    assert err < tol, err


def test_unreachable_modes_raise():
    from vface_b200.ldm.models.pnp_utils import register_spa_attn_injection
    h, _, _ = _build(torch.float32)
    for bad in ("temporal", "adaIn", "mix", "fft_vfixed"):
        with pytest.raises(NotImplementedError):
            register_spa_attn_injection(h, 1, input_blocks=True, fusion=bad)
    with pytest.raises(NotImplementedError):
        register_spa_attn_injection(h, 1, input_blocks=True, chunks=2)
