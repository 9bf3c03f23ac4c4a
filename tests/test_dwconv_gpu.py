"""nz_dwconv3x3_fwd / _bwd (csrc/dwconv_kernels.cu) against the reference's own expression
``act(conv2d(x))`` (nnunetv2/nets/m2net.py:214-215: depthwise Conv2d 3x3 pad 1 + SiLU) evaluated in fp64 with autograd.
Floating-point kernel: rel 1e-3 (max-norm) fp32, 2e-2 for 16-bit I/O (north_star tolerances)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [(2, 32, 64, 64), (1, 64, 17, 33), (3, 8, 1, 1), (1, 4, 2, 130), (2, 16, 9, 5), (1, 256, 16, 16), (1, 32, 512, 512)]


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
@pytest.mark.parametrize("has_bias", [True, False])
def test_dwconv_silu_matches_reference_expression(shape, dtype, has_bias):
    from nnuzoo_b200.dwconv import dwconv3x3_silu
    B, D, H, W = shape
    if H * W > 100000 and (dtype != torch.bfloat16 or not has_bias):
        pytest.skip("full-size plane once")
    torch.manual_seed(H * 1000 + W)
    x = torch.randn(*shape, device="cuda").to(dtype).requires_grad_(True)
    w = (0.4 * torch.randn(D, 1, 3, 3, device="cuda")).requires_grad_(True)
    b = (0.2 * torch.randn(D, device="cuda")).requires_grad_(True) if has_bias else None
    gy = torch.randn(*shape, device="cuda").to(dtype)
    y = dwconv3x3_silu(x, w, b)
    y.backward(gy)
    x64 = x.detach().double().requires_grad_(True)
    w64 = w.detach().double().requires_grad_(True)
    b64 = b.detach().double().requires_grad_(True) if has_bias else None
    y64 = F.silu(F.conv2d(x64, w64, b64, padding=1, groups=D))
    y64.backward(gy.double())
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert y.dtype == dtype and _rel(y, y64.detach()) < tol
    assert _rel(x.grad, x64.grad) < tol
    assert _rel(w.grad, w64.grad) < (1e-3 if dtype == torch.float32 else tol)
    if has_bias:
        assert _rel(b.grad, b64.grad) < (1e-3 if dtype == torch.float32 else tol)


def test_plain_convolution_mode_and_cpu_refusal():
    from nnuzoo_b200.dwconv import dwconv3x3_silu
    x = torch.randn(1, 4, 6, 7, device="cuda")
    w = torch.randn(4, 1, 3, 3, device="cuda")
    assert torch.allclose(dwconv3x3_silu(x, w, None, silu=False), F.conv2d(x, w, None, padding=1, groups=4), atol=1e-5)
    with pytest.raises(RuntimeError):
        dwconv3x3_silu(x.cpu(), w.cpu(), None)
