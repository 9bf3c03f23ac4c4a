"""The fused 1-D inner node (nnuzoo_b200/mamba_inner.py: conv -> x_proj -> dt_proj -> scan [-> out_proj], conv / delta
recomputed in the backward) against the op-by-op statement of MambaInnerFn.forward
(selective_scan_interface.py:292-367) built from the individually parity-tested ops, and the reversed direction against
the flipped call it replaces (mamba_simple.py:250-262).  The reference-generated goldens for the whole block
(tests/golden/module_mamba_*.npz, test_module_gpu.py) run through the same node.  fp32: rel 1e-3 (scan tolerance);
bf16 under autocast: 2e-2."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _params(d_model, d_inner, R, N, W, seed, dev):
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, sc=1.0: (sc * torch.randn(*s, generator=g)).to(dev).requires_grad_(True)  # noqa: E731
    return dict(conv_w=rn(d_inner, 1, W, sc=0.4), conv_b=rn(d_inner, sc=0.1), x_w=rn(R + 2 * N, d_inner, sc=d_inner ** -0.5),
                dt_w=rn(d_inner, R, sc=R ** -0.5), out_w=rn(d_model, d_inner, sc=d_inner ** -0.5),
                A_log=torch.log(torch.arange(1, N + 1).float()).repeat(d_inner, 1).to(dev).requires_grad_(True),
                D=rn(d_inner), bias=rn(d_inner, sc=0.5))


def _unfused(xz, p, R, N, out_proj=True):
    """MambaInnerFn.forward op for op (selective_scan_interface.py:315-367) on the stand-alone ops."""
    from nnuzoo_b200 import causal_conv1d_fn, selective_scan_fn
    x, z = xz.chunk(2, dim=1)
    x = causal_conv1d_fn(x, p["conv_w"].squeeze(1), p["conv_b"], "silu")
    x_dbl = F.linear(x.transpose(1, 2), p["x_w"])
    dt, B, C = torch.split(x_dbl, [R, N, N], dim=-1)
    delta = F.linear(dt, p["dt_w"]).transpose(1, 2).contiguous()
    B = B.transpose(1, 2).unsqueeze(1).contiguous()
    C = C.transpose(1, 2).unsqueeze(1).contiguous()
    y = selective_scan_fn(x, delta, -torch.exp(p["A_log"]), B, C, p["D"], z=z, delta_bias=p["bias"], delta_softplus=True)
    return F.linear(y.transpose(1, 2), p["out_w"]) if out_proj else y


def _fused(xz, p, out_proj=True, reverse=False):
    from nnuzoo_b200.mamba_inner import mamba_inner_fn, mamba_inner_fn_no_out_proj
    A = -torch.exp(p["A_log"])
    if out_proj:
        return mamba_inner_fn(xz, p["conv_w"], p["conv_b"], p["x_w"], p["dt_w"], p["out_w"], None, A, None, None, p["D"],
                              delta_bias=p["bias"], delta_softplus=True, reverse=reverse)
    return mamba_inner_fn_no_out_proj(xz, p["conv_w"], p["conv_b"], p["x_w"], p["dt_w"], A, None, None, p["D"],
                                      delta_bias=p["bias"], delta_softplus=True, reverse=reverse)


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


def _grads(fn, xz, p, gy):
    for t in [xz, *p.values()]:
        t.grad = None
    y = fn()
    y.backward(gy)
    return y.detach(), {"xz": xz.grad.clone(), **{k: v.grad.clone() for k, v in p.items() if v.grad is not None}}


# (batch, d_model, L, out_proj): L = 75 / 600 are MambaND's token counts (mamba_nd2net.py:972-994), 4096 takes the
# row-per-lane kernels when reversed, 37 is ragged (scalar conv path)
@pytest.mark.parametrize("shape", [(2, 16, 75, True), (2, 24, 600, True), (1, 16, 37, False), (2, 32, 4096, True),
                                   (1, 16, 2048, False)])
def test_fused_inner_matches_op_by_op(shape):
    batch, d_model, L, out_proj = shape
    dev = torch.device("cuda:0")
    d_inner, R, N, W = 2 * d_model, max(1, d_model // 16), 16, 4
    p = _params(d_model, d_inner, R, N, W, 11, dev)
    g = torch.Generator().manual_seed(5)
    xz = torch.randn(batch, 2 * d_inner, L, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(*((batch, L, d_model) if out_proj else (batch, d_inner, L)), generator=g).to(dev)
    y0, g0 = _grads(lambda: _unfused(xz, p, R, N, out_proj), xz, p, gy)
    y1, g1 = _grads(lambda: _fused(xz, p, out_proj), xz, p, gy)
    assert _rel(y1, y0) < 1e-3
    for k in g0:
        if k == "out_w" and not out_proj:
            continue
        assert _rel(g1[k], g0[k]) < 2e-3, k


@pytest.mark.parametrize("shape", [(2, 16, 75, True), (2, 32, 4096, True), (1, 16, 2048, False), (2, 16, 8192, True)])
def test_reversed_direction_equals_flipped_call(shape):
    """reverse=True == fn(xz.flip(-1)).flip(L) (mamba_simple.py:250-262); at L = 2048 / 4096 / 8192 with 32+ rows per
    group it runs by addressing (anti-causal conv + reversed-walk scan), at L = 75 through the flip fallback."""
    batch, d_model, L, out_proj = shape
    dev = torch.device("cuda:0")
    d_inner, R, N, W = 2 * d_model, max(1, d_model // 16), 16, 4
    p = _params(d_model, d_inner, R, N, W, 13, dev)
    g = torch.Generator().manual_seed(6)
    xz = torch.randn(batch, 2 * d_inner, L, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(*((batch, L, d_model) if out_proj else (batch, d_inner, L)), generator=g).to(dev)
    tdim = 1 if out_proj else -1
    y0, g0 = _grads(lambda: _fused(xz.flip(-1), p, out_proj).flip(tdim), xz, p, gy)
    y1, g1 = _grads(lambda: _fused(xz, p, out_proj, reverse=True), xz, p, gy)
    assert _rel(y1, y0) < 1e-3
    for k in g0:
        assert _rel(g1[k], g0[k]) < 2e-3, k


def test_reverse_by_addressing_is_taken_when_expressible():
    from nnuzoo_b200.mamba_inner import _rev_by_addressing
    dev = torch.device("cuda:0")
    xz = torch.empty(2, 128, 4096, device=dev)
    x, z = xz.chunk(2, dim=1)
    assert _rev_by_addressing(x, z, 16, 2)
    xz = torch.empty(2, 128, 75, device=dev)
    x, z = xz.chunk(2, dim=1)
    assert not _rev_by_addressing(x, z, 16, 2)


def test_fused_inner_bf16_autocast():
    batch, d_model, L = 2, 32, 4096
    dev = torch.device("cuda:0")
    d_inner, R, N, W = 2 * d_model, 2, 16, 4
    p = _params(d_model, d_inner, R, N, W, 17, dev)
    g = torch.Generator().manual_seed(8)
    xz = torch.randn(batch, 2 * d_inner, L, generator=g).to(dev).requires_grad_(True)
    gy = torch.randn(batch, L, d_model, generator=g).to(dev)
    y0, g0 = _grads(lambda: _fused(xz, p), xz, p, gy)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y1, g1 = _grads(lambda: _fused(xz, p).float(), xz, p, gy)
    assert _rel(y1, y0) < 2e-2
    for k in ("xz", "x_w", "dt_w", "out_w", "conv_w", "D", "bias", "A_log"):
        assert _rel(g1[k].float(), g0[k]) < 5e-2, k


def test_mamba_module_reverse_flag():
    """Mamba(...)(h, reverse=True) == Mamba(...)(h.flip(1)).flip(1) -- MambaND's reversed layers (mamba_nd2net.py:638-656)."""
    from nnuzoo_b200 import Mamba
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    for d_model, L in ((16, 75), (32, 2048)):
        m = Mamba(d_model=d_model).to(dev)
        h = torch.randn(2, L, d_model, device=dev, requires_grad=True)
        gy = torch.randn(2, L, d_model, device=dev)
        y0 = m(h.flip(1)).flip(1)
        y0.backward(gy)
        gh0, gp0 = h.grad.clone(), {k: q.grad.clone() for k, q in m.named_parameters() if q.grad is not None}
        h.grad = None
        m.zero_grad()
        y1 = m(h, reverse=True)
        y1.backward(gy)
        assert _rel(y1, y0) < 1e-3 and _rel(h.grad, gh0) < 2e-3
        for k, q in m.named_parameters():
            if q.grad is not None:
                assert _rel(q.grad, gp0[k]) < 2e-3, k
