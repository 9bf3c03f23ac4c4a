"""torch.compile over the scan: the reference compiles its networks by default (nnUNetTrainer.py:316-322), so the scan must
be one traceable custom op (nnuzoo_b200/library_ops.py), not a graph break; and CUDA-graph capture of the C-ABI calls
("must be CUDA-graph-capturable")."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _inputs(batch, dim, L, groups=1, z=True, dtype=torch.float32, seed=0):
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(seed)
    r = lambda *s: torch.randn(*s, device=dev, generator=g).to(dtype)  # noqa: E731
    u, delta = r(batch, dim, L).requires_grad_(True), (0.5 * r(batch, dim, L)).requires_grad_(True)
    A = (-torch.arange(1, 17, device=dev).float().repeat(dim, 1)).requires_grad_(True)
    B, C = r(batch, groups, 16, L).requires_grad_(True), r(batch, groups, 16, L).requires_grad_(True)
    D = torch.ones(dim, device=dev, requires_grad=True)
    zz = r(batch, dim, L).requires_grad_(True) if z else None
    bias = torch.full((dim,), -2.0, device=dev, requires_grad=True)
    return [u, delta, A, B, C, D, zz, bias], r(batch, dim, L)


@pytest.mark.parametrize("shape", [(2, 64, 600, 1), (2, 128, 4096, 4)])
def test_compiled_scan_matches_eager_without_graph_breaks(shape):
    from nnuzoo_b200 import selective_scan_fn
    batch, dim, L, groups = shape
    leaves, gout = _inputs(batch, dim, L, groups)

    def f(u, delta, A, B, C, D, z, bias):
        return selective_scan_fn(u * 1.0, delta, A, B, C, D, z, bias, True) * 2.0

    ref = f(*leaves)
    ref.backward(gout)
    g_ref = [t.grad.clone() for t in leaves if t is not None]
    for t in leaves:
        if t is not None:
            t.grad = None
    cf = torch.compile(f, fullgraph=True)        # fullgraph: a graph break at the scan would raise here
    out = cf(*leaves)
    out.backward(gout)
    g = [t.grad for t in leaves if t is not None]
    torch.cuda.synchronize()
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    for a, b in zip(g, g_ref):
        assert torch.allclose(a, b, rtol=1e-4, atol=1e-5)


def test_opcheck_of_the_registered_ops():
    from nnuzoo_b200 import library_ops  # noqa: F401
    leaves, _ = _inputs(1, 32, 256, 1, z=False)
    u, delta, A, B, C, D, _, bias = [t.detach() if t is not None else None for t in leaves]
    torch.library.opcheck(torch.ops.nnuzoo_b200.scan_fwd.default,
                          (u, delta, A, B, C, D, None, bias, True, False, False, 0, 1),
                          test_utils=("test_schema", "test_faketensor"))


def test_scan_fwd_bwd_is_cuda_graph_capturable():
    """A MambaND-sized scan (BASELINE configs[3]) captured once and replayed: same numbers as the eager call."""
    from nnuzoo_b200 import selective_scan_fn
    leaves, gout = _inputs(2, 192, 600, 1)
    torch.cuda.synchronize()

    def step():
        for t in leaves:
            t.grad = None
        o = selective_scan_fn(*leaves[:8], True)
        o.backward(gout)
        return o

    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            ref = step()
    torch.cuda.current_stream().wait_stream(s)
    g_ref = [t.grad.clone() for t in leaves]
    ref = ref.detach().clone()
    graph = torch.cuda.CUDAGraph()
    for t in leaves:
        t.grad = None
    with torch.cuda.graph(graph):
        out = step()
    for _ in range(3):
        graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out.detach(), ref)
    for t, r in zip(leaves, g_ref):
        assert torch.allclose(t.grad, r, rtol=1e-5, atol=1e-6)    # (dA etc. accumulate with atomics: order may differ)
