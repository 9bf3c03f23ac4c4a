"""nz_layernorm_fwd / _bwd (csrc/norm_kernels.cu) against torch's fp64 layer_norm and its autograd on the same
inputs.  Floating-point kernel: rel 1e-3 (max-norm) for fp32 outputs, 2e-2 for 16-bit (north_star tolerances);
statistics are fp32 in every mode."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

SHAPES = [((3, 37, 16), 16), ((2, 64, 64, 32), 32), ((1000, 64), 64), ((5, 128), 128), ((7, 3, 256), 256),
          ((33, 512), 512), ((9, 1024), 1024), ((1, 1, 16), 16)]
DTYPES = [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16), (torch.float16, torch.float16),
          (torch.bfloat16, torch.float32), (torch.float32, torch.bfloat16), (torch.float32, torch.float16)]


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("shape,C", SHAPES)
@pytest.mark.parametrize("din,dout", DTYPES)
def test_layernorm_matches_torch_fp64(shape, C, din, dout):
    from nnuzoo_b200.norm import LayerNormFn
    torch.manual_seed(C + len(shape))
    x = (torch.randn(*shape, device="cuda") * 2 + 0.5).to(din).requires_grad_(True)
    w = (1 + 0.2 * torch.randn(C, device="cuda")).requires_grad_(True)
    b = (0.1 * torch.randn(C, device="cuda")).requires_grad_(True)
    gy = torch.randn(*shape, device="cuda").to(dout)
    y = LayerNormFn.apply(x, w, b, 1e-5, dout)
    y.backward(gy)
    x64 = x.detach().double().requires_grad_(True)
    w64, b64 = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    y64 = F.layer_norm(x64, (C,), w64, b64, 1e-5)
    y64.backward(gy.double())
    tol_out = 1e-3 if dout == torch.float32 else 2e-2
    tol_in = 1e-3 if din == torch.float32 else 2e-2
    assert y.dtype == dout and _rel(y, y64.detach()) < tol_out
    assert x.grad.dtype == din and _rel(x.grad, x64.grad) < max(tol_in, tol_out if din != torch.float32 else 1e-3)
    assert _rel(w.grad, w64.grad) < 1e-3 and _rel(b.grad, b64.grad) < 1e-3


def test_module_is_a_drop_in_and_follows_autocast_rule():
    from nnuzoo_b200.norm import LayerNorm
    torch.manual_seed(0)
    ours, ref = LayerNorm(64).cuda(), torch.nn.LayerNorm(64).cuda()
    ours.load_state_dict(ref.state_dict(), strict=True)
    x = torch.randn(4, 50, 64, device="cuda")
    assert torch.allclose(ours(x), ref(x), atol=2e-6, rtol=1e-5)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        xb = x.bfloat16()
        a, b = ours(xb), ref(xb)
        assert a.dtype == b.dtype == torch.float32 and torch.allclose(a, b, atol=2e-6, rtol=1e-5)
        ours.feeds_linear = True
        c = ours(x)
        want = ref(x).bfloat16()                                  # fp32 LayerNorm, then the Linear's cast
        assert c.dtype == torch.bfloat16 and (c == want).float().mean() > 0.999   # (fp32 summation order may
        assert float((c.float() - want.float()).abs().max()) <= 2 ** -6           #  flip a last bit here and there)
    odd = LayerNorm(96).cuda()                                                   # 96 / 4 = 24: not a power of two
    assert odd(torch.randn(3, 96, device="cuda")).shape == (3, 96)               # library route, same semantics
