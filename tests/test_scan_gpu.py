"""GPU parity tests for the selective scan (run on a B200: ``pytest -m gpu``).

The CUDA path (through the C ABI) is compared with
  * the committed golden vectors produced by the reference itself (tests/golden/scan_*.npz), and
  * the CPU oracle (oracle/scan_oracle.c, fp64 flavour) on seeded inputs at sizes it finishes in
    seconds, including BASELINE config 1 and the long-L stage-1 shape,
with the tolerances BASELINE.json's north_star states: rel 1e-3 in fp32, 2e-2 in bf16
(rel = max|a-b| / max|b| per tensor).
"""
import numpy as np
import pytest
import torch

from tests.helpers import golden_names, load_golden, rel_err, scan_inputs

pytestmark = pytest.mark.gpu

TOL = {"float32": 1e-3, "bfloat16": 2e-2, "float16": 2e-2}
GRADS = {"du": "u", "ddelta": "delta", "dA": "A", "dB": "B", "dC": "C", "dD": "D", "dz": "z",
         "ddelta_bias": "delta_bias"}


def _dev():
    if not torch.cuda.is_available():
        pytest.fail("GPU tests need a CUDA device (there is no CPU fallback for this path)")
    return torch.device("cuda:0")


def _to_dev(inp, dt, requires_grad=True):
    dtype = getattr(torch, dt)
    out = {}
    for k, v in inp.items():
        if v is None:
            out[k] = None
            continue
        t = torch.from_numpy(np.ascontiguousarray(v)).to(_dev())
        if k in ("u", "delta", "B", "C", "z"):
            t = t.to(dtype)
        out[k] = t.requires_grad_(requires_grad)
    return out


def _run(inp, softplus, gout, return_last_state=True, force_generic=False):
    import nnuzoo_b200.selective_scan_interface as ssi
    ssi._FORCE_GENERIC = force_generic
    try:
        out, last = ssi.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                                          inp["delta_bias"], softplus, True)
        out.backward(gout)
    finally:
        ssi._FORCE_GENERIC = False
    torch.cuda.synchronize()
    grads = {g: (inp[k].grad if inp[k] is not None else None) for g, k in GRADS.items()}
    return out, last, grads


@pytest.mark.parametrize("name", golden_names("scan_"))
def test_golden_forward_backward(name):
    rec = load_golden(name)
    inp_np, softplus, dt = scan_inputs(rec)
    inp = _to_dev(inp_np, dt)
    gout = torch.from_numpy(rec["in_gout"]).to(_dev()).to(getattr(torch, dt))
    out, last, grads = _run(inp, softplus, gout)
    assert out.dtype == getattr(torch, dt) and tuple(out.shape) == rec["out"].shape
    assert rel_err(out.detach().float().cpu().numpy(), rec["out"]) < TOL[dt]
    assert rel_err(last.detach().cpu().numpy(), rec["last_state"]) < TOL[dt]
    for g, k in GRADS.items():
        if "grad_" + k in rec:
            got = grads[g].float().cpu().numpy()
            assert got.shape == rec["grad_" + k].shape, g
            assert rel_err(got, rec["grad_" + k]) < TOL[dt], (g, rel_err(got, rec["grad_" + k]))
        else:
            assert grads[g] is None


def _seeded(batch, dim, groups, N, L, has_z, seed, dt="float32", split_views=False):
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(batch, dim, L, generator=g)
    delta = 0.5 * torch.randn(batch, dim, L, generator=g)
    A = -(torch.arange(1, N + 1).float().repeat(dim, 1) * torch.exp(0.1 * torch.randn(dim, N, generator=g)))
    if split_views:
        # what SS2D hands over: B, C are split views of one (b, k, R + 2N, L) projection (m2net.py:181)
        R = 6
        xdbl = torch.randn(batch, groups, R + 2 * N, L, generator=g)
        B, C = xdbl[:, :, R:R + N], xdbl[:, :, R + N:]
    else:
        B = torch.randn(batch, groups, N, L, generator=g)
        C = torch.randn(batch, groups, N, L, generator=g)
    D = 1.0 + 0.1 * torch.randn(dim, generator=g)
    dtv = torch.exp(torch.rand(dim, generator=g) * (np.log(0.1) - np.log(0.001)) + np.log(0.001))
    bias = dtv + torch.log(-torch.expm1(-dtv))
    z = torch.randn(batch, dim, L, generator=g) if has_z else None
    gout = torch.randn(batch, dim, L, generator=g)
    dtype = getattr(torch, dt)
    cast = lambda t: None if t is None else t.to(dtype)  # noqa: E731
    return dict(u=cast(u), delta=cast(delta), A=A, B=cast(B), C=cast(C), D=D, z=cast(z), delta_bias=bias), cast(gout)


def _oracle(inp, gout, softplus=True):
    from oracle import scan_oracle
    f = {k: (None if v is None else v.float()) for k, v in inp.items()}
    out, last = scan_oracle.selective_scan_oracle(**{k: f[k] for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias")},
                                                  delta_softplus=softplus, return_last_state=True, precision="f64")
    grads = scan_oracle.selective_scan_oracle_bwd(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], f["z"],
                                                  f["delta_bias"], softplus, gout.float(), precision="f64")
    return out, last, grads


def _compare(inp_cpu, gout_cpu, dt, force_generic=False, keep_views=False):
    dev = _dev()
    inp = {}
    for k, v in inp_cpu.items():
        if v is None:
            inp[k] = None
        elif keep_views and k in ("B", "C"):
            inp[k] = v  # filled below from the shared parent
        else:
            inp[k] = v.to(dev).requires_grad_(True)
    if keep_views:
        # rebuild the parent buffer on the device so B and C stay non-contiguous split views
        parent = inp_cpu["B"]._base.to(dev).requires_grad_(True)
        R = parent.shape[2] - 2 * inp_cpu["B"].shape[2]
        N = inp_cpu["B"].shape[2]
        inp["B"], inp["C"] = parent[:, :, R:R + N], parent[:, :, R + N:]
        assert not inp["B"].is_contiguous()
    import nnuzoo_b200.selective_scan_interface as ssi
    ssi._FORCE_GENERIC = force_generic
    try:
        out, last = ssi.selective_scan_fn(inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"], inp["z"],
                                          inp["delta_bias"], True, True)
        out.backward(gout_cpu.to(dev))
    finally:
        ssi._FORCE_GENERIC = False
    torch.cuda.synchronize()
    ref_out, ref_last, ref_g = _oracle(inp_cpu, gout_cpu)
    tol = TOL[dt]
    errs = {"out": rel_err(out.detach().float().cpu().numpy(), ref_out), "last": rel_err(last.detach().cpu().numpy(), ref_last)}
    if keep_views:
        R = parent.shape[2] - 2 * N
        pg = parent.grad.float().cpu().numpy()
        errs["dB"] = rel_err(pg[:, :, R:R + N], ref_g["dB"])
        errs["dC"] = rel_err(pg[:, :, R + N:], ref_g["dC"])
    for g, k in GRADS.items():
        if ref_g[g] is None or (keep_views and k in ("B", "C")):
            continue
        errs[g] = rel_err(inp[k].grad.float().cpu().numpy(), ref_g[g])
    bad = {k: v for k, v in errs.items() if not (v < tol)}
    assert not bad, f"rel errors above {tol}: {bad} (all: {errs})"
    return errs


# (batch, dim, groups, N, L, has_z): TMA-eligible and not, grouped and not, ragged tails, tiny
SHAPES = [
    (2, 32, 4, 16, 256, False),     # one chunk, 8 rows per group
    (2, 64, 2, 16, 1024, True),     # z gate, 4 chunks, 32 rows per group
    (1, 16, 1, 16, 4096 + 128, False),  # L not a multiple of the chunk (TMA zero-fills the tail)
    (2, 8, 1, 16, 75, True),        # MambaND2Net token count: generic path, ragged
    (1, 24, 3, 16, 600, False),     # generic path (L % 32 != 0), 8 rows per group
    (2, 6, 2, 16, 320, False),      # 3 rows per group -> one-row-per-CTA kernel
    (1, 8, 1, 8, 512, False),       # d_state 8
    (1, 8, 1, 16, 1, False),        # L = 1
    (2, 64, 4, 16, 512, False),     # exactly one 16-row tile per group: TMA path, plain dB/dC stores
    (1, 96, 2, 16, 384, True),      # 48 rows per group: three row blocks share dB/dC (vector atomics), z gate
    (3, 40, 2, 16, 640, False),     # 20 rows per group: generic path with a partly filled second row block
]


@pytest.mark.parametrize("shape", SHAPES)
def test_seeded_vs_oracle_fp32(shape):
    batch, dim, groups, N, L, has_z = shape
    inp, gout = _seeded(batch, dim, groups, N, L, has_z, seed=hash(shape) % 1000)
    _compare(inp, gout, "float32")


@pytest.mark.parametrize("shape", SHAPES[:3])
def test_generic_loader_matches_oracle_on_tma_shapes(shape):
    batch, dim, groups, N, L, has_z = shape
    inp, gout = _seeded(batch, dim, groups, N, L, has_z, seed=7)
    _compare(inp, gout, "float32", force_generic=True)


@pytest.mark.parametrize("dt", ["bfloat16", "float16"])
@pytest.mark.parametrize("L", [512, 200])
def test_16bit_io(dt, L):
    inp, gout = _seeded(2, 16, 1, 16, L, True, seed=11, dt=dt)
    _compare(inp, gout, dt)


def test_noncontiguous_split_views_as_ss2d_passes_them():
    """SURVEY.md hard part 5: B/C arrive as torch.split views; strides go to the kernel as they are."""
    inp, gout = _seeded(2, 32, 4, 16, 1024, False, seed=3, split_views=True)
    _compare(inp, gout, "float32", keep_views=True)


def test_baseline_config1_shape():
    """BASELINE.json configs[0]: B=2, K=4, d_inner=192, d_state=16, L=64*64, fp32."""
    inp, gout = _seeded(2, 768, 4, 16, 4096, False, seed=0, split_views=True)
    _compare(inp, gout, "float32", keep_views=True)


def test_long_sequence_stage1_shape():
    """M2Net stage-1 scan at inference batch 1: K*D = 128, L = 512*512 (tolerance on the LONGEST L)."""
    inp, gout = _seeded(1, 128, 4, 16, 512 * 512, False, seed=5)
    _compare(inp, gout, "float32")


def test_linearity_in_u_at_full_size():
    """Size-independent property: with z=None the op is linear in u for fixed delta, B, C."""
    dev = _dev()
    from nnuzoo_b200 import selective_scan_fn
    g = torch.Generator(device=dev).manual_seed(0)
    batch, dim, G, N, L = 4, 128, 4, 16, 512 * 512
    u1 = torch.randn(batch, dim, L, device=dev, generator=g)
    u2 = torch.randn(batch, dim, L, device=dev, generator=g)
    delta = 0.5 * torch.randn(batch, dim, L, device=dev, generator=g)
    A = -torch.arange(1, N + 1, device=dev).float().repeat(dim, 1)
    B = torch.randn(batch, G, N, L, device=dev, generator=g)
    C = torch.randn(batch, G, N, L, device=dev, generator=g)
    D = torch.ones(dim, device=dev)
    bias = torch.full((dim,), -3.0, device=dev)
    f = lambda u: selective_scan_fn(u, delta, A, B, C, D, None, bias, True)  # noqa: E731
    lhs = f(u1 + u2)
    rhs = f(u1) + f(u2)
    assert torch.isfinite(lhs).all()
    assert float((lhs - rhs).abs().max()) <= 1e-3 * float(rhs.abs().max())


def test_deterministic_forward():
    inp, _ = _seeded(2, 64, 2, 16, 2048, True, seed=9)
    dev = _dev()
    from nnuzoo_b200 import selective_scan_fn
    args = [None if v is None else v.to(dev) for v in (inp["u"], inp["delta"], inp["A"], inp["B"], inp["C"], inp["D"],
                                                       inp["z"], inp["delta_bias"])]
    a = selective_scan_fn(*args, True)
    b = selective_scan_fn(*args, True)
    assert torch.equal(a, b)


def test_rejects_what_it_does_not_implement():
    dev = _dev()
    from nnuzoo_b200 import selective_scan_fn
    u = torch.zeros(1, 8, 16, device=dev)
    with pytest.raises(NotImplementedError):
        selective_scan_fn(u, u, torch.zeros(8, 32, device=dev), torch.zeros(1, 1, 32, 16, device=dev),
                          torch.zeros(1, 1, 32, 16, device=dev))
    with pytest.raises(RuntimeError):
        selective_scan_fn(u.cpu(), u.cpu(), torch.zeros(8, 16), torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))


def test_host_entry_point_matches_device_path():
    """nz_scan_fwd_bwd_host (host pointers, pipelined over batch slices) == the device-pointer path."""
    import ctypes

    from nnuzoo_b200 import _native
    lib = _native.lib()
    _native.bind_device(0)
    batch, dim, G, N, L = 3, 64, 2, 16, 768
    inp, gout = _seeded(batch, dim, G, N, L, True, seed=7)
    dev_inp = {k: (None if v is None else v.to(_dev()).requires_grad_(True)) for k, v in inp.items()}
    out, last, grads = _run(dev_inp, True, gout.to(_dev()))
    f = lambda a: np.ascontiguousarray(a.detach().cpu().numpy() if hasattr(a, "detach") else a, dtype=np.float32)  # noqa: E731
    h = {k: f(v) for k, v in inp.items() if v is not None}
    res = {k: np.empty_like(h["u"]) for k in ("out", "du", "ddelta", "dz")}
    res.update(dB=np.empty_like(h["B"]), dC=np.empty_like(h["C"]), dA=np.empty_like(h["A"]),
               dD=np.empty_like(h["D"]), db=np.empty_like(h["delta_bias"]))
    hg = f(gout)
    p = lambda a: ctypes.c_void_p(a.ctypes.data)  # noqa: E731
    d = _native.NzScanDesc()
    d.batch, d.dim, d.dstate, d.ngroups, d.seqlen = batch, dim, N, G, L
    d.dtype, d.delta_softplus = 0, 1
    d.u, d.delta, d.A, d.B, d.C, d.D, d.z, d.delta_bias = (p(h[k]) for k in ("u", "delta", "A", "B", "C", "D", "z", "delta_bias"))
    for st in (d.u_stride, d.delta_stride, d.z_stride, d.out_stride, d.dout_stride):
        st[0], st[1] = dim * L, L
    for st in (d.B_stride, d.C_stride):
        st[0], st[1], st[2] = G * N * L, N * L, L
    d.A_stride = N
    d.out, d.dout = p(res["out"]), p(hg)
    d.du, d.ddelta, d.dz = p(res["du"]), p(res["ddelta"]), p(res["dz"])
    d.dA, d.dB, d.dC, d.dD, d.ddelta_bias = p(res["dA"]), p(res["dB"]), p(res["dC"]), p(res["dD"]), p(res["db"])
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _native.check(lib.nz_scan_fwd_bwd_host(ctypes.byref(d), stream), "nz_scan_fwd_bwd_host")
    pairs = {"out": out, "du": grads["du"], "ddelta": grads["ddelta"], "dz": grads["dz"], "dB": grads["dB"],
             "dC": grads["dC"], "dA": grads["dA"], "dD": grads["dD"], "db": grads["ddelta_bias"]}
    for k, ref in pairs.items():
        assert rel_err(res[k], ref.detach().float().cpu().numpy()) < 1e-5, k


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
def test_16bit_operands_fp32_result_equals_widened_fp32_scan(dtype):
    """out_dtype=float32 extension (SS2D under autocast): 16-bit operands in, fp32 out.  Same operand values, same
    fp32 arithmetic as the reference's widen-then-scan (m2net.py:185-200): outputs agree to fp32 rounding; gradients
    to the 16-bit tolerance (dout is read, du / ddelta are written in the operand dtype)."""
    from nnuzoo_b200 import selective_scan_fn
    torch.manual_seed(5)
    Bn, K, D, N, L = 2, 4, 32, 16, 1536
    mk = lambda *s: torch.randn(*s, device="cuda")  # noqa: E731
    u16 = mk(Bn, K * D, L).to(dtype)
    dl16 = (0.5 * mk(Bn, K * D, L)).to(dtype)
    xdbl = mk(Bn, K, 2 + 2 * N, L).to(dtype)
    B16, C16 = xdbl[:, :, 2:2 + N], xdbl[:, :, 2 + N:]          # strided split views, as SS2D hands them over
    A = -torch.exp(torch.log(torch.arange(1, N + 1, device="cuda").float()).repeat(K * D, 1) + 0.1 * mk(K * D, N))
    Dp, bias = 1 + 0.1 * mk(K * D), -2 + 0.3 * mk(K * D)
    gout = mk(Bn, K * D, L)

    def run(u, dl, Bm, Cm, **kw):
        leaves = [t.detach().clone().requires_grad_(True) for t in (u, dl, Bm, Cm)]
        p = [t.detach().clone().requires_grad_(True) for t in (A, Dp, bias)]
        out = selective_scan_fn(leaves[0], leaves[1], p[0], leaves[2], leaves[3], p[1], None, p[2], True, **kw)
        out.backward(gout.to(out.dtype))
        return out.detach(), [t.grad for t in leaves + p]

    o_mix, g_mix = run(u16, dl16, B16, C16, out_dtype=torch.float32)
    o_ref, g_ref = run(u16.float(), dl16.float(), B16.float(), C16.float())
    assert o_mix.dtype == torch.float32
    assert rel_err(o_mix.cpu().numpy(), o_ref.cpu().numpy()) < 1e-5
    for a, b in zip(g_mix, g_ref):
        assert rel_err(a.float().cpu().numpy(), b.float().cpu().numpy()) < 2e-2


@pytest.mark.gpu
@pytest.mark.parametrize("case", [(1, 64, 1, 10240, torch.bfloat16, True), (2, 32, 2, 4196, torch.float32, False),
                                  (1, 128, 4, 65536, torch.float32, False), (2, 48, 1, 5000, torch.float16, True)])
def test_chunk_parallel_forward_is_bit_identical_to_the_chained_one(case, monkeypatch):
    """Few row blocks + many chunks: nz_scan_fwd runs aggregate pass / combine / final pass (capi.cu run_fwd_cp) when the
    workspace has the _cp size.  The combine performs the chained kernel's own fma per link, so out, the checkpoints
    (hence every gradient) and last_state must be IDENTICAL to the chained run (NZ_NO_CP=1)."""
    from nnuzoo_b200 import _native, selective_scan_fn
    Bn, D, G, L, dtype, has_z = case
    torch.manual_seed(L)
    mk = lambda *s: torch.randn(*s, device="cuda")  # noqa: E731
    u, dl = mk(Bn, D, L).to(dtype), (0.5 * mk(Bn, D, L)).to(dtype)
    z = mk(Bn, D, L).to(dtype) if has_z else None
    Bm, Cm = mk(Bn, G, 16, L).to(dtype), mk(Bn, G, 16, L).to(dtype)
    A = -torch.exp(0.3 * mk(D, 16)) * torch.arange(1, 17, device="cuda")
    Dp, bias = mk(D), -2 + 0.3 * mk(D)
    gout = mk(Bn, D, L).to(dtype)

    def run():
        leaves = [t.detach().clone().requires_grad_(True) for t in (u, dl, Bm, Cm, A, Dp, bias)]
        n0 = _native.launch_count()
        out, last = selective_scan_fn(leaves[0], leaves[1], leaves[4], leaves[2], leaves[3], leaves[5], z, leaves[6],
                                      True, True)
        launches = _native.launch_count() - n0
        out.backward(gout)
        return out.detach(), last.detach(), [t.grad for t in leaves], launches

    o_cp, l_cp, g_cp, n_cp = run()
    monkeypatch.setenv("NZ_NO_CP", "1")
    o_ch, l_ch, g_ch, n_ch = run()
    assert (n_cp, n_ch) == (3, 1), "the chunk-parallel path must actually have been taken"
    assert torch.equal(o_cp, o_ch) and torch.equal(l_cp, l_ch)
    for a, b in zip(g_cp, g_ch):   # same checkpoints -> same gradients up to the run-to-run order of the fp32 atomics
        tol = 1e-5 if a.dtype == torch.float32 else 1e-2     # (one flipped last bit after rounding to 16 bits)
        assert rel_err(a.float().cpu().numpy(), b.float().cpu().numpy()) < tol
