"""Module-level parity: our SS2D / SSND against the reference modules' recorded behaviour.

tests/golden/module_*.npz hold state_dict + input + output + gradients of the REFERENCE modules
(m2net.py:39-225, ssnd2net.py:73-318, seg_mamba/mamba_simple.py:37-357) run with the reference's own
selective_scan_ref.  The GEMMs
around the scan run in a different order on the GPU, so the tolerance is the scan's (rel 1e-3).
"""
import numpy as np
import pytest
import torch

from tests.helpers import golden_names, load_golden, rel_err

pytestmark = pytest.mark.gpu


def _build(name, rec):
    from nnuzoo_b200 import SS2D, SSND, Mamba
    d_model = rec["x"].shape[-1]
    if name.startswith("module_mamba"):
        mod = Mamba(d_model=d_model, bimamba_type=name.rsplit("_", 1)[1], nslices=5)
    elif name.startswith("module_ss2d"):
        mod = SS2D(d_model=d_model)
    elif "ssnd2d" in name:
        mod = SSND(spatial_dims=2, factorization_type="cross-scan", d_model=d_model)
    else:
        mod = SSND(spatial_dims=3, factorization_type="cross-scan", d_model=d_model)
    sd = {k[3:]: torch.from_numpy(v) for k, v in rec.items() if k.startswith("sd_")}
    mod.load_state_dict(sd, strict=True)  # reference checkpoints load unchanged
    return mod.cuda().eval()


@pytest.mark.parametrize("name", [n for n in golden_names("module_") if "m2net" not in n])
def test_module_forward_backward_matches_reference(name):
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    rec = load_golden(name)
    mod = _build(name, rec)
    x = torch.from_numpy(rec["x"]).cuda().requires_grad_(True)
    y = mod(x)
    assert tuple(y.shape) == rec["y"].shape
    assert rel_err(y.detach().cpu().numpy(), rec["y"]) < 1e-3
    y.backward(torch.from_numpy(rec["gy"]).cuda())
    assert rel_err(x.grad.cpu().numpy(), rec["gx"]) < 1e-3
    for k, p in mod.named_parameters():
        want = rec["gp_" + k]
        got = np.zeros_like(want) if p.grad is None else p.grad.cpu().numpy()
        scale = max(float(np.abs(want).max()), 1e-6)
        assert float(np.abs(got - want).max()) / scale < 2e-3, k
