"""GPU parity for CrossScan / CrossMerge: bit-exact against the reference fixtures and the oracle."""
import numpy as np
import pytest
import torch

from oracle import cross_oracle
from tests.helpers import golden_names, load_golden

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device")
    return torch.device("cuda:0")


@pytest.mark.parametrize("name", golden_names("cross_"))
def test_golden_bit_exact(name):
    from nnuzoo_b200 import cross_merge, cross_scan
    rec = load_golden(name)
    x = torch.from_numpy(rec["x"]).to(_dev())
    xs = cross_scan(x)
    assert np.array_equal(xs.cpu().numpy(), rec["xs"])
    y = cross_merge(torch.from_numpy(rec["out_y"]).to(_dev()), x.shape[2:], "reference")
    assert np.array_equal(y.cpu().numpy(), rec["y"])


@pytest.mark.parametrize("shape", [(2, 16, 64, 64), (1, 8, 17, 33), (1, 4, 8, 12, 10), (2, 3, 5, 7, 3)])
@pytest.mark.parametrize("dt", ["float32", "bfloat16"])
def test_random_shapes_bit_exact_vs_oracle(shape, dt):
    from nnuzoo_b200 import cross_merge, cross_scan
    g = torch.Generator().manual_seed(len(shape) * 100 + shape[-1])
    x = torch.randn(*shape, generator=g).to(getattr(torch, dt))
    xs = cross_scan(x.to(_dev()))
    scan = cross_oracle.cross_scan_2d if len(shape) == 4 else cross_oracle.cross_scan_3d
    want = scan(x.float().numpy())
    assert np.array_equal(xs.float().cpu().numpy(), want)
    if dt == "float32":
        K = xs.shape[1]
        oy = torch.randn(shape[0], K, shape[1], int(np.prod(shape[2:])), generator=g)
        for mode in (["reference"] if len(shape) == 4 else ["reference", "fixed"]):
            got = cross_merge(oy.to(_dev()), shape[2:], mode).cpu().numpy()
            if len(shape) == 4:
                ref = cross_oracle.cross_merge_2d(oy.numpy(), *shape[2:])
            else:
                ref = cross_oracle.cross_merge_3d(oy.numpy(), *shape[2:], mode=mode)
            assert np.array_equal(got, ref), mode


@pytest.mark.parametrize("shape", [(1, 4, 6, 5), (1, 2, 3, 4, 5)])
def test_adjoints_are_exact_transposes(shape):
    """<cross_scan(x), g> == <x, cross_scan^T(g)> and the same for cross_merge (both modes in 3-D)."""
    from nnuzoo_b200 import cross_merge, cross_scan
    dev = _dev()
    g = torch.Generator().manual_seed(1)
    x = torch.randn(*shape, generator=g, dtype=torch.float64).float().to(dev).requires_grad_(True)
    xs = cross_scan(x)
    gxs = torch.randn(xs.shape, generator=g).to(dev)
    xs.backward(gxs)
    lhs = float((xs.detach().double() * gxs.double()).sum())
    rhs = float((x.detach().double() * x.grad.double()).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs))
    for mode in (["reference"] if len(shape) == 4 else ["reference", "fixed"]):
        oy = torch.randn(xs.shape, generator=g).to(dev).requires_grad_(True)
        y = cross_merge(oy, shape[2:], mode)
        gy = torch.randn(y.shape, generator=g).to(dev)
        y.backward(gy)
        lhs = float((y.detach().double() * gy.double()).sum())
        rhs = float((oy.detach().double() * oy.grad.double()).sum())
        assert abs(lhs - rhs) <= 1e-4 * max(1.0, abs(lhs)), mode
        if mode == "reference" and len(shape) == 5:
            assert float(oy.grad[:, 2].abs().max()) == 0.0 and float(oy.grad[:, 5].abs().max()) == 0.0


def test_roundtrip_property_full_size():
    """merge(scan(x)) == ((x + x) + x) + x exactly, at the M2Net stage-1 size (batch 2)."""
    from nnuzoo_b200 import cross_merge, cross_scan
    dev = _dev()
    x = torch.randn(2, 32, 512, 512, device=dev)
    y = cross_merge(cross_scan(x), (512, 512)).view_as(x)
    assert torch.equal(y, ((x + x) + x) + x)


@pytest.mark.parametrize("shape", [(2, 8, 24, 40, 36), (1, 5, 7, 33, 65)])
def test_tiled_3d_equals_generic_kernels(shape, monkeypatch):
    """The tiled 3-D passes (re-factored (Z*H) x W, Z x (H*W), H x (W*Z) matrices through 32 x 32 tiles) against the
    one-thread-per-element kernels they replace (NZ_CROSS_GENERIC=1), bit for bit: scan, merge (both modes), merge
    adjoint -- at sizes with ragged tiles."""
    from nnuzoo_b200 import cross_merge, cross_scan
    dev = _dev()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(*shape, generator=g).to(dev)
    L = int(np.prod(shape[2:]))
    oy = torch.randn(shape[0], 6, shape[1], L, generator=g).to(dev)
    gy = torch.randn(shape[0], shape[1], L, generator=g).to(dev)

    def run():
        res = [cross_scan(x), cross_scan(x.to(torch.bfloat16))]
        for mode in ("reference", "fixed"):
            o = oy.clone().requires_grad_(True)
            y = cross_merge(o, shape[2:], mode)
            y.backward(gy.view_as(y))
            res += [y.detach(), o.grad]
        return res

    tiled = run()
    monkeypatch.setenv("NZ_CROSS_GENERIC", "1")
    generic = run()
    for a, b in zip(tiled, generic):
        assert torch.equal(a, b)
