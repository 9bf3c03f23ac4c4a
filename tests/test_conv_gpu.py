"""GPU parity of the depthwise causal conv1d (+ SiLU) against the CPU oracle (oracle/conv_oracle.py, which is
pinned to the reference expression mamba_simple.py:316-317 in tests/test_oracle_golden.py).
Tolerances: rel 1e-5 in fp32 (same arithmetic, different summation order), 2e-2 with 16-bit I/O."""
import numpy as np
import pytest
import torch

from tests.helpers import rel_err

pytestmark = pytest.mark.gpu

# (batch, dim, L, width, dtype, bias, silu, view): view = x is a channel-chunk view of a (b, 2*dim, L) tensor
CASES = [
    (2, 8, 64, 4, "float32", True, True, False),
    (2, 6, 37, 4, "float32", True, True, False),      # ragged L: scalar path
    (1, 4, 1, 4, "float32", True, True, False),       # L = 1
    (3, 16, 2048 + 4, 3, "float32", False, True, False),
    (2, 8, 96, 2, "float32", True, False, False),     # no activation
    (2, 8, 128, 4, "float32", True, True, True),      # xz.chunk view as the Mamba block passes it
    (2, 8, 256, 4, "bfloat16", True, True, True),
    (2, 8, 250, 4, "float16", True, True, False),
]


@pytest.mark.parametrize("case", CASES)
def test_causal_conv1d_matches_oracle(case):
    from nnuzoo_b200 import causal_conv1d_fn
    from oracle import conv_oracle
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    batch, dim, L, W, dt, has_bias, silu, view = case
    dtype = getattr(torch, dt)
    g = torch.Generator().manual_seed(hash(case) % 1000)
    full = torch.randn(batch, 2 * dim, L, generator=g).to(dtype)
    w = torch.randn(dim, W, generator=g)
    b = torch.randn(dim, generator=g) if has_bias else None
    go = torch.randn(batch, dim, L, generator=g).to(dtype)
    dev = torch.device("cuda:0")
    fd = full.to(dev).requires_grad_(True)
    xd = fd[:, :dim] if view else fd[:, :dim].contiguous()
    wd = w.to(dev).requires_grad_(True)
    bd = None if b is None else b.to(dev).requires_grad_(True)
    out = causal_conv1d_fn(xd, wd, bd, "silu" if silu else None)
    out.backward(go.to(dev))
    torch.cuda.synchronize()
    xs = full[:, :dim].float().numpy()
    ref = conv_oracle.causal_conv1d_oracle(xs, w.numpy(), None if b is None else b.numpy(), silu)
    rg = conv_oracle.causal_conv1d_oracle_bwd(xs, w.numpy(), None if b is None else b.numpy(), go.float().numpy(), silu)
    tol = 1e-5 if dt == "float32" else 2e-2
    assert rel_err(out.detach().float().cpu().numpy(), ref) < tol
    assert rel_err(fd.grad[:, :dim].float().cpu().numpy(), rg["dx"]) < tol
    assert float(fd.grad[:, dim:].abs().max()) == 0.0
    assert rel_err(wd.grad.cpu().numpy(), rg["dw"]) < max(tol, 1e-4)
    if b is not None:
        assert rel_err(bd.grad.cpu().numpy(), rg["dbias"]) < max(tol, 1e-4)


@pytest.mark.parametrize("case", CASES)
def test_reverse_conv1d_is_the_flipped_convolution(case):
    """reverse=True == causal_conv1d_fn(x.flip(-1)).flip(-1) (mamba_simple.py:250-262) without the flipped copies:
    checked against the same oracle applied to the flipped sequence."""
    from nnuzoo_b200 import causal_conv1d_fn
    from oracle import conv_oracle
    batch, dim, L, W, dt, has_bias, silu, view = case
    dtype = getattr(torch, dt)
    g = torch.Generator().manual_seed(7 + hash(case) % 1000)
    full = torch.randn(batch, 2 * dim, L, generator=g).to(dtype)
    w = torch.randn(dim, W, generator=g)
    b = torch.randn(dim, generator=g) if has_bias else None
    go = torch.randn(batch, dim, L, generator=g).to(dtype)
    dev = torch.device("cuda:0")
    fd = full.to(dev).requires_grad_(True)
    xd = fd[:, :dim] if view else fd[:, :dim].contiguous()
    wd = w.to(dev).requires_grad_(True)
    bd = None if b is None else b.to(dev).requires_grad_(True)
    out = causal_conv1d_fn(xd, wd, bd, "silu" if silu else None, reverse=True)
    out.backward(go.to(dev))
    torch.cuda.synchronize()
    xs = np.ascontiguousarray(full[:, :dim].float().numpy()[..., ::-1])
    gs = np.ascontiguousarray(go.float().numpy()[..., ::-1])
    bn = None if b is None else b.numpy()
    ref = conv_oracle.causal_conv1d_oracle(xs, w.numpy(), bn, silu)[..., ::-1]
    rg = conv_oracle.causal_conv1d_oracle_bwd(xs, w.numpy(), bn, gs, silu)
    tol = 1e-5 if dt == "float32" else 2e-2
    assert rel_err(out.detach().float().cpu().numpy(), ref) < tol
    assert rel_err(fd.grad[:, :dim].float().cpu().numpy(), rg["dx"][..., ::-1]) < tol
    assert rel_err(wd.grad.cpu().numpy(), rg["dw"]) < max(tol, 1e-4)
    if b is not None:
        assert rel_err(bd.grad.cpu().numpy(), rg["dbias"]) < max(tol, 1e-4)
