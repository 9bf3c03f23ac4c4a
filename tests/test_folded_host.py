"""Host logic of the folded SS2D core (nnuzoo_b200/ss2d.py::_folded_core) on the CPU: the permutation of the per-direction
parameters into the folded direction order (k' = 2 * array + backwards), the stacked x_proj weights per array, and the
merge association -- against the reference SS2D fixtures (tests/golden/module_ss2d_*.npz).

THIS TEST replaces the CUDA pieces by CPU stand-ins (pair CrossScan = two reshapes, grouped projections = einsum, the fused
scan + merge + LayerNorm + gate node = the oracle scan on explicitly flipped copies followed by torch ops).  The GPU suite
runs the same fixtures through the kernels (tests/test_module_gpu.py, tests/test_folded_gpu.py)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle.torch_port import selective_scan_port
from tests.helpers import load_golden, rel_err


def _pair(x):
    b, d, h, w = x.shape
    return torch.stack([x.reshape(b, d, h * w), x.transpose(2, 3).reshape(b, d, h * w)], 1)


def _proj(x, w):
    return torch.einsum("bknl,kmn->bkml", x, w).contiguous()


def _core(xs2, dts, As, Bs, Cs, Ds, bias, z, gamma, beta, H, W, eps, out_dtype):
    bsz, _, D, L = xs2.shape
    ys = []
    for k in range(4):                                  # folded order: k' = 2 * array + backwards
        a, rev = k // 2, k % 2
        f = (lambda t: t.flip(-1)) if rev else (lambda t: t)
        sl = slice(k * D, (k + 1) * D)
        o = selective_scan_port(f(xs2[:, a]), f(dts[:, k]), As[sl], f(Bs[:, k]), f(Cs[:, k]), Ds[sl], z=None,
                                delta_bias=bias[sl], delta_softplus=True)
        ys.append(f(o))                                 # back at un-flipped positions
    t = lambda y: y.view(bsz, D, W, H).transpose(2, 3).reshape(bsz, D, L)   # noqa: E731  column-major -> row-major
    y = ((ys[0] + ys[1]) + t(ys[2])) + t(ys[3])         # m2net.py:218 association: y0 + inv_y0 + wh_y + invwh_y
    y = F.layer_norm(y.transpose(1, 2).reshape(bsz, H, W, D), (D,), gamma, beta, eps)
    return y * F.silu(z)


@pytest.mark.parametrize("name", ["module_ss2d_m8", "module_ss2d_m32"])
def test_folded_core_host_logic_matches_reference_ss2d(name, monkeypatch):
    import nnuzoo_b200.ss2d as ss
    monkeypatch.setattr(ss, "cross_scan_pair", _pair)
    monkeypatch.setattr(ss, "grouped_proj", _proj)
    monkeypatch.setattr(ss.fused, "ss2d_core_folded", _core)
    rec = load_golden(name)
    m = ss.SS2D(d_model=rec["x"].shape[-1])
    m.load_state_dict({k[3:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("sd_")}, strict=True)
    m.eval()
    m.fuse_projections = False
    x = torch.from_numpy(rec["x"]).requires_grad_(True)
    # SS2D.forward (m2net.py:208-225) with the folded core called directly (its CUDA gate is what is stood in for)
    xc, z = m.in_proj(x).chunk(2, dim=-1)
    xc = m.act(m.conv2d(xc.permute(0, 3, 1, 2).contiguous()))
    y = m.out_proj(m._folded_core(xc, z))
    assert rel_err(y.detach().numpy(), rec["y"]) < 1e-4
    y.backward(torch.from_numpy(rec["gy"]))
    assert rel_err(x.grad.numpy(), rec["gx"]) < 1e-4
    for k, p in m.named_parameters():
        want = rec["gp_" + k]
        got = np.zeros_like(want) if p.grad is None else p.grad.numpy()
        scale = max(float(np.abs(want).max()), 1e-6)
        assert float(np.abs(got - want).max()) / scale < 1e-3, k
