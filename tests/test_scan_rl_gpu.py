"""GPU parity tests of the row-per-lane kernels (csrc/scan_rl_kernels.cuh): the chunk-parallel backward fed by the fine
checkpoints, and the row-per-lane forward.

By default nz_scan_* only picks them from ~25 M elements up (capi.cu rl_shape_ok); NZ_RL_MIN_ELTS=0 makes the small
shapes the fp64 oracle can check take the same code.  Tolerances as everywhere: rel 1e-3 fp32, 2e-2 16-bit.
"""
import ctypes

import numpy as np
import pytest
import torch

from tests.helpers import rel_err
from tests.test_scan_gpu import _compare, _seeded

pytestmark = pytest.mark.gpu


def _launches_of_backward(inp, gout):
    """Run fwd + bwd of selective_scan_fn on the GPU; return (out, grads, launches of fwd, launches of bwd)."""
    from nnuzoo_b200 import _native, selective_scan_fn
    dev = torch.device("cuda:0")
    leaves = {k: (None if v is None else v.to(dev).requires_grad_(True)) for k, v in inp.items()}
    n0 = _native.launch_count()
    out = selective_scan_fn(leaves["u"], leaves["delta"], leaves["A"], leaves["B"], leaves["C"], leaves["D"], leaves["z"],
                            leaves["delta_bias"], True)
    n1 = _native.launch_count()
    out.backward(gout.to(dev))
    n2 = _native.launch_count()
    torch.cuda.synchronize()
    return out.detach(), {k: v.grad for k, v in leaves.items() if v is not None}, n1 - n0, n2 - n1


# (batch, dim, groups, N, L, has_z)
RL_SHAPES = [
    (2, 128, 4, 16, 1024, False),   # 32 rows per group: one warp owns each dB / dC element (plain stores)
    (2, 128, 2, 16, 4128, False),   # 64 rows per group: two warps RED into dB / dC; L not a multiple of the forward tile
    (2, 64, 1, 16, 2048, True),     # z gate, one group (the 1-D nets' layout)
    (1, 96, 1, 16, 992, True),      # 96 rows per group, 31 tiles: uneven chunks
    (3, 32, 1, 16, 32, False),      # a single 32-step tile
]


@pytest.mark.parametrize("fwd_rl", ["0", "1"])
@pytest.mark.parametrize("shape", RL_SHAPES)
def test_row_per_lane_vs_oracle_fp32(shape, fwd_rl, monkeypatch):
    monkeypatch.setenv("NZ_RL_MIN_ELTS", "0")
    monkeypatch.setenv("NZ_RL_FWD", fwd_rl)
    batch, dim, groups, N, L, has_z = shape
    inp, gout = _seeded(batch, dim, groups, N, L, has_z, seed=13 + L)
    _compare(inp, gout, "float32")


@pytest.mark.parametrize("items", ["1", "7", "100000"])
def test_chunk_count_does_not_change_the_result(items, monkeypatch):
    """One chunk per row block (no aggregate pass), a few, and one chunk per 128-byte tile."""
    monkeypatch.setenv("NZ_RL_MIN_ELTS", "0")
    monkeypatch.setenv("NZ_RL_ITEMS", items)
    inp, gout = _seeded(2, 64, 2, 16, 1536, False, seed=21)
    _compare(inp, gout, "float32")
    _, _, nf, nb = _launches_of_backward(inp, gout)
    assert nb == (1 if items == "1" else 3), "aggregate pass + combine + main pass, or the main pass alone"


@pytest.mark.parametrize("dt", ["bfloat16", "float16"])
@pytest.mark.parametrize("fwd_rl", ["0", "1"])
def test_row_per_lane_16bit_io(dt, fwd_rl, monkeypatch):
    monkeypatch.setenv("NZ_RL_MIN_ELTS", "0")
    monkeypatch.setenv("NZ_RL_FWD", fwd_rl)
    inp, gout = _seeded(2, 64, 2, 16, 2048, True, seed=17, dt=dt)
    _compare(inp, gout, dt)


def test_row_per_lane_takes_split_views(monkeypatch):
    """B / C as the strided split views SS2D hands over (no .contiguous() copy on this path either)."""
    monkeypatch.setenv("NZ_RL_MIN_ELTS", "0")
    inp, gout = _seeded(2, 128, 4, 16, 1024, False, seed=3, split_views=True)
    _compare(inp, gout, "float32", keep_views=True)


def test_default_dispatch_large_shape_agrees_with_warp_scan(monkeypatch):
    """33.5 M elements: the default dispatch takes the row-per-lane backward (3 launches); its results agree with the
    warp-scan kernels (NZ_NO_RL=1) far inside the tolerance, and with the fp64 oracle through them (test_scan_gpu)."""
    inp, gout = _seeded(4, 128, 4, 16, 65536, False, seed=2)
    out_rl, g_rl, nf, nb = _launches_of_backward(inp, gout)
    assert nb == 3, "row-per-lane backward expected for this size"
    monkeypatch.setenv("NZ_NO_RL", "1")
    out_ws, g_ws, _, nb_ws = _launches_of_backward(inp, gout)
    assert nb_ws == 1
    assert rel_err(out_rl.cpu().numpy(), out_ws.cpu().numpy()) < 1e-5
    for k in g_rl:
        assert rel_err(g_rl[k].float().cpu().numpy(), g_ws[k].float().cpu().numpy()) < 1e-4, k


def test_inference_does_not_take_fine_checkpoints(monkeypatch):
    """No gradient wanted -> no xf buffer, one forward launch path as before."""
    from nnuzoo_b200 import selective_scan_fn
    monkeypatch.setenv("NZ_RL_MIN_ELTS", "0")
    inp, _ = _seeded(2, 128, 4, 16, 1024, False, seed=4)
    dev = torch.device("cuda:0")
    args = [None if v is None else v.to(dev) for v in (inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"],
                                                       inp["z"], inp["delta_bias"])]
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    out = selective_scan_fn(*args, True)
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    fine = 2 * 128 * (1024 // 8) * 16 * 4
    assert peak < out.numel() * 4 + fine, "the forward must not have allocated the fine checkpoints"


@pytest.mark.parametrize("shape", [(2, 128, 2, 16, 4128, False), (1, 96, 1, 16, 992, True), (2, 256, 2, 16, 2048, False)])
@pytest.mark.parametrize("items", ["7", "100000"])
def test_backward_zero_fills_db_dc_itself(shape, items, monkeypatch):
    """With several row blocks per group the row-per-lane backward accumulates dB / dC with RED; its aggregate pass
    zero-fills them first (RlArgs::zero_dbc, nz_scan_bwd_overwrites_dbc() == 1), so the caller hands over uninitialised
    buffers.  Every torch.empty of this test returns NaN-poisoned memory: any element the kernels fail to write or to
    zero shows up in the comparison with the fp64 oracle."""
    from nnuzoo_b200 import _native
    from nnuzoo_b200._native import NzScanDesc
    monkeypatch.setenv("NZ_RL_MIN_ELTS", "0")
    monkeypatch.setenv("NZ_RL_ITEMS", items)
    batch, dim, groups, N, L, has_z = shape
    d = NzScanDesc()
    d.batch, d.dim, d.dstate, d.ngroups, d.seqlen, d.dtype = batch, dim, N, groups, L, 0
    d.u = d.delta = d.B = d.C = ctypes.c_void_p(4096)
    d.xf = ctypes.c_void_p(4096)
    for s in (d.u_stride, d.delta_stride):
        s[0], s[1] = dim * L, L
    for s in (d.B_stride, d.C_stride):
        s[0], s[1], s[2] = groups * N * L, N * L, L
    if items == "100000":  # many chunks: the aggregate pass exists and does the zero fill (one chunk: the caller zeroes)
        assert _native.lib().nz_scan_bwd_overwrites_dbc(ctypes.byref(d)) == 1
    orig = torch.empty

    def poisoned(*a, **k):
        t = orig(*a, **k)
        return t.fill_(float("nan")) if t.is_floating_point() and t.is_cuda else t

    monkeypatch.setattr(torch, "empty", poisoned)
    inp, gout = _seeded(batch, dim, groups, N, L, has_z, seed=5 + L)
    _compare(inp, gout, "float32")
