"""The callers of the 1-D scan path against the reference's own classes (tests/golden/shell_*.npz, made by
oracle/gen_golden.py:gen_shells from lm2net.py:64-176 and mamba_nd2net.py:565-666, :725-1001 running the reference's
selective_scan_ref): MambaLayer, ResMambaBlock with its axis orders, MambaND's Block for every order x direction, and
MambaNDCore stacks.  Reference state_dicts load (the reference's vendored Mamba also carries unused ``*_b`` / ``*_s``
parameters, which upstream ``mamba_ssm.Mamba`` -- the class these nets import -- does not have; they are dropped).
Tolerance: the scan's (rel 1e-3 forward / input gradient, 2e-3 of the largest entry for parameter gradients)."""
import json
import os

import numpy as np
import pytest
import torch

from tests.helpers import GOLDEN, golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu



@pytest.fixture(autouse=True)
def _exact_fp32_library_convs():
    """The GSC / patch-embedding convolutions are cuDNN's; keep them in true fp32 so the comparison measures our path."""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


_MAN = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))["module"]


def _unused(k):
    return any(t in k for t in ("_b.", "_s.", "A_b_log", "A_s_log", "D_b", "D_s"))


def _build(name, rec):
    from nnuzoo_b200.mamba_nd import MambaLayer, MambaNDCore, ResMambaBlock, create_block
    meta = _MAN[name]
    if name == "shell_mambalayer":
        mod, call = MambaLayer(input_dim=16, output_dim=24), lambda m, x: m(x)
    elif name.startswith("shell_resmamba"):
        mod = ResMambaBlock(meta["spatial_dims"], 16, norm=("GROUP", {"num_groups": 8}), order=meta["order"])
        call = lambda m, x: m(x)  # noqa: E731
    elif name.startswith("shell_ndblock"):
        mod = create_block(spatial_dims=3, d_model=16, ssm_cfg={"d_state": 16}, fused_add_norm=False, residual_in_fp32=True,
                           reverse=meta["reverse"], drop_rate=0.0, drop_path_rate=0.0)
        call = lambda m, x: m(x, order=meta["order"], shape=tuple(meta["shape"]), n_dim_pos=4)  # noqa: E731
    elif name == "shell_ndcore":
        mod = MambaNDCore(spatial_dims=3, img_size=(4, 6, 6), patch_size=(2, 2, 2), in_channels=2, embed_dims=16,
                          num_layers=7, fused_add_norm=False, final_norm=False)
        call = lambda m, x: m(x)[0]  # noqa: E731
    else:
        mod = MambaNDCore(spatial_dims=2, img_size=(8, 8), patch_size=(2, 2), in_channels=1, embed_dims=16, num_layers=4,
                          fused_add_norm=False, final_norm=False)
        call = lambda m, x: m(x)[0]  # noqa: E731
    sd = {k[3:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("sd_") and not _unused(k[3:])}
    mod.load_state_dict(sd, strict=True)
    return mod.cuda().eval(), call


@pytest.mark.parametrize("name", golden_names("shell_"))
def test_shell_matches_reference(name):
    rec = load_golden(name)
    mod, call = _build(name, rec)
    x = torch.from_numpy(rec["x"]).cuda().requires_grad_(True)
    y = call(mod, x)
    assert tuple(y.shape) == rec["y"].shape
    assert rel_err(y.detach().cpu().numpy(), rec["y"]) < 1e-3
    y.backward(torch.from_numpy(rec["gy"]).cuda())
    assert rel_err(x.grad.cpu().numpy(), rec["gx"]) < 1e-3
    for k, p in mod.named_parameters():
        want = rec["gp_" + k]
        got = np.zeros_like(want) if p.grad is None else p.grad.cpu().numpy()
        scale = max(float(np.abs(want).max()), 1e-6)
        assert float(np.abs(got - want).max()) / scale < 2e-3, k


def test_reversed_block_at_a_size_the_reversed_walk_takes():
    """A reversed Block on 4096 tokens x 32 channels runs by addressing (anti-causal conv + reversed-walk scan).  It must
    equal: order the tokens, flip, forward block in identity order, flip, un-order (mamba_nd2net.py:631-662)."""
    from nnuzoo_b200.mamba_nd import _token_perm, create_block
    torch.manual_seed(1)
    rev = create_block(3, 32, ssm_cfg={"d_state": 16}, reverse=True, drop_rate=0.0, drop_path_rate=0.0).cuda()
    fwd = create_block(3, 32, ssm_cfg={"d_state": 16}, reverse=False, drop_rate=0.0, drop_path_rate=0.0).cuda()
    fwd.load_state_dict(rev.state_dict())
    n, t, h, w, c = 2, 16, 8, 32, 32
    x = torch.randn(n, t * h * w, c, device="cuda")
    for order in ("t h w", "t w h", "w h t"):
        perm = _token_perm(order, 4)
        inv = [perm.index(i) for i in range(5)]
        dims = [(n, t, h, w, c)[p] for p in perm]
        ordered = x.view(n, t, h, w, c).permute(perm).reshape(n, -1, c)
        y0 = fwd(ordered.flip(1), order="t h w", shape=(t, h, w)).flip(1)
        y0 = y0.reshape(dims).permute(inv).reshape(n, -1, c)
        y1 = rev(x, order=order, shape=(t, h, w))
        assert rel_err(y1.detach().cpu().numpy(), y0.detach().cpu().numpy()) < 1e-3, order
