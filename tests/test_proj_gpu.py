"""grouped_proj / nz_proj_wgrad (csrc/proj_kernels.cu) against the reference's own expression -- the einsums at
nnunetv2/nets/m2net.py:179 and :182 and their autograd -- evaluated in fp64 on the same inputs.
Floating-point kernel: tolerance rel 1e-3 (max-norm) for fp32 I/O, 2e-2 for bf16 / fp16 (north_star)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

# (B, K, N, M, L): x_proj of the four M2Net stages, dt_proj of the same, ragged / tiny sequences
CASES = [(2, 4, 32, 33, 4096), (2, 4, 64, 34, 1024), (1, 4, 128, 36, 640), (1, 4, 256, 40, 256),
         (2, 4, 1, 32, 4096), (2, 4, 8, 256, 512), (3, 6, 16, 34, 75), (1, 4, 32, 33, 1), (2, 4, 32, 33, 8191)]


def _ref(x, w, g):
    x64 = x.double().requires_grad_(True)
    w64 = w.double().requires_grad_(True)
    y = torch.einsum("b k n l, k m n -> b k m l", x64, w64)
    y.backward(g.double())
    return y.detach(), x64.grad, w64.grad


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("case", CASES)
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16, torch.float16])
def test_grouped_proj_matches_einsum(case, dtype):
    from nnuzoo_b200.proj import grouped_proj
    B, K, N, M, L = case
    torch.manual_seed(B * 1000 + L)
    x = torch.randn(B, K, N, L, device="cuda").to(dtype).requires_grad_(True)
    w = (torch.randn(K, M, N, device="cuda") / N ** 0.5).requires_grad_(True)
    g = torch.randn(B, K, M, L, device="cuda").to(dtype)
    y = grouped_proj(x, w)
    y.backward(g)
    wq = w.detach().to(dtype)          # the forward / dx GEMMs run in the I/O dtype, as the reference's do
    y_ref, dx_ref, dw_ref = _ref(x.detach(), wq, g)
    tol = 1e-3 if dtype == torch.float32 else 2e-2
    assert y.dtype == dtype and _rel(y, y_ref) < tol
    assert _rel(x.grad, dx_ref) < tol
    assert w.grad.dtype == torch.float32 and _rel(w.grad, dw_ref) < (1e-4 if dtype == torch.float32 else tol)


def test_wgrad_strided_split_view_and_mixed_dtypes():
    """dt_proj's X is the (B, K, R, L) split view of x_dbl (row stride L, direction stride C*L); G may be fp32."""
    from nnuzoo_b200.proj import proj_wgrad
    B, K, C, R, D, L = 2, 4, 36, 4, 128, 1000
    torch.manual_seed(3)
    x_dbl = torch.randn(B, K, C, L, device="cuda").bfloat16()
    xr = x_dbl[:, :, :R]
    assert not xr.is_contiguous()
    g = torch.randn(B, K, D, L, device="cuda")
    dw = proj_wgrad(g, xr)
    ref = torch.einsum("b k m l, b k n l -> k m n", g.double(), xr.double())
    assert _rel(dw, ref) < 1e-4


def test_wgrad_is_the_sum_of_its_parts():
    """Linearity in the position range: full-size stage-1 shape split in two halves along L."""
    from nnuzoo_b200.proj import proj_wgrad
    B, K, M, N, L = 2, 4, 33, 32, 262144
    torch.manual_seed(4)
    g = torch.randn(B, K, M, L, device="cuda").bfloat16()
    x = torch.randn(B, K, N, L, device="cuda").bfloat16()
    full = proj_wgrad(g, x)
    halves = proj_wgrad(g[..., :L // 2], x[..., :L // 2]) + proj_wgrad(g[..., L // 2:], x[..., L // 2:])
    assert _rel(full, halves.double()) < 1e-4


def test_cpu_tensors_are_refused():
    from nnuzoo_b200.proj import proj_wgrad
    with pytest.raises(RuntimeError):
        proj_wgrad(torch.zeros(1, 1, 2, 4), torch.zeros(1, 1, 2, 4))


def test_wide_projection_falls_back_to_the_library_gemm():
    """M * N above the kernel's accumulator budget (d_model 192: (R + 2N) x d_inner = 44 x 384) must not raise."""
    from nnuzoo_b200.proj import grouped_proj
    torch.manual_seed(0)
    x = torch.randn(2, 4, 384, 96, device="cuda", requires_grad=True)
    w = torch.randn(4, 44, 384, device="cuda", requires_grad=True)
    y = grouped_proj(x, w)
    g = torch.randn_like(y)
    y.backward(g)
    ref = torch.einsum("bkml,bknl->kmn", g.double(), x.detach().double())
    assert float((w.grad.double() - ref).abs().max() / ref.abs().max()) < 1e-4
