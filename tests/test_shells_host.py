"""Host logic of the 1-D callers (nnuzoo_b200/mamba.py, mamba_nd.py) on the CPU: token orders, the no-flip treatment of
reversed Blocks, GSC / MambaLayer wiring and state_dict naming -- against the reference-made fixtures
(tests/golden/shell_*.npz, module_mamba_*.npz).

The product's CUDA ops cannot run here, so THIS TEST swaps the two fused entry points for CPU stand-ins built from the
oracle (oracle/torch_port.py's restated selective_scan_ref + F.conv1d; a reversed direction = flip, run, flip back).  The
oracle stays test infrastructure: nothing in the package imports it, and the GPU suite (tests/test_shells_gpu.py) runs the
same fixtures through the real kernels."""
import json
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.torch_port import selective_scan_port
from tests.helpers import GOLDEN, golden_names, load_golden, rel_err

_MAN = json.load(open(os.path.join(GOLDEN, "MANIFEST.json")))["module"]


def _inner_cpu(xz, conv_w, conv_b, x_w, dt_w, A, D, bias, reverse):
    """MambaInnerFnNoOutProj.forward (selective_scan_interface.py:159-226) with CPU ops."""
    if reverse:
        return _inner_cpu(xz.flip(-1), conv_w, conv_b, x_w, dt_w, A, D, bias, False).flip(-1)
    L = xz.shape[-1]
    R, N = dt_w.shape[1], A.shape[1]
    x, z = xz.chunk(2, dim=1)
    w = conv_w.reshape(conv_w.shape[0], 1, conv_w.shape[-1])
    x = F.silu(F.conv1d(x, w, conv_b, padding=w.shape[-1] - 1, groups=x.shape[1])[..., :L])
    x_dbl = torch.matmul(x_w, x)
    delta = torch.matmul(dt_w, x_dbl[:, :R])
    B = x_dbl[:, R:R + N].unsqueeze(1)
    C = x_dbl[:, R + N:].unsqueeze(1)
    return selective_scan_port(x, delta, A, B, C, D, z=z, delta_bias=bias, delta_softplus=True)


@pytest.fixture
def cpu_inner(monkeypatch):
    import nnuzoo_b200.mamba as mm

    def no_out_proj(xz, conv_w, conv_b, x_w, dt_w, A, B=None, C=None, D=None, delta_bias=None, B_proj_bias=None,
                    C_proj_bias=None, delta_softplus=True, *, reverse=False):
        return _inner_cpu(xz, conv_w, conv_b, x_w, dt_w, A, D, delta_bias, reverse)

    def with_out_proj(xz, conv_w, conv_b, x_w, dt_w, out_w, out_b, A, B=None, C=None, D=None, delta_bias=None,
                      B_proj_bias=None, C_proj_bias=None, delta_softplus=True, *, reverse=False):
        y = _inner_cpu(xz, conv_w, conv_b, x_w, dt_w, A, D, delta_bias, reverse)
        return F.linear(y.transpose(1, 2), out_w, out_b)

    monkeypatch.setattr(mm, "mamba_inner_fn", with_out_proj)
    monkeypatch.setattr(mm, "mamba_inner_fn_no_out_proj", no_out_proj)


def _unused(k):
    return any(t in k for t in ("_b.", "_s.", "A_b_log", "A_s_log", "D_b", "D_s"))


def _build(name, rec):
    from nnuzoo_b200.mamba import Mamba
    from nnuzoo_b200.mamba_nd import MambaLayer, MambaNDCore, ResMambaBlock, create_block
    meta = _MAN[name]
    drop_unused = True
    if name.startswith("module_mamba"):
        mod = Mamba(d_model=rec["x"].shape[-1], bimamba_type=meta["bimamba_type"], nslices=5)
        call, drop_unused = (lambda m, x: m(x)), False
    elif name == "shell_mambalayer":
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
    sd = {k[3:]: torch.from_numpy(v) for k, v in rec.items()
          if k.startswith("sd_") and not (drop_unused and _unused(k[3:]))}
    mod.load_state_dict(sd, strict=True)
    return mod.eval(), call


@pytest.mark.parametrize("name", golden_names("shell_") + golden_names("module_mamba"))
def test_host_logic_matches_reference_fixtures(name, cpu_inner):
    rec = load_golden(name)
    mod, call = _build(name, rec)
    x = torch.from_numpy(rec["x"]).requires_grad_(True)
    y = call(mod, x)
    assert tuple(y.shape) == rec["y"].shape
    assert rel_err(y.detach().numpy(), rec["y"]) < 1e-4
    y.backward(torch.from_numpy(rec["gy"]))
    assert rel_err(x.grad.numpy(), rec["gx"]) < 1e-4
    for k, p in mod.named_parameters():
        want = rec["gp_" + k]
        got = np.zeros_like(want) if p.grad is None else p.grad.numpy()
        scale = max(float(np.abs(want).max()), 1e-6)
        assert float(np.abs(got - want).max()) / scale < 1e-3, k
