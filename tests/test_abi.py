"""CPU-side checks of the C-ABI boundary and the host layer (no GPU, no compute calls)."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "nnuzoo_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(nz_[a-z0-9_]+)\s*\(", hdr)))


def test_library_loads_and_exports_every_declared_symbol():
    from nnuzoo_b200 import _native
    from nnuzoo_b200.build import build_native
    build_native()
    lib = ctypes.CDLL(_native.LIB_PATH)
    syms = _declared_symbols()
    assert {"nz_scan_fwd", "nz_scan_bwd", "nz_cross_scan", "nz_cross_merge", "nz_cross_merge_bwd",
            "nz_scan_fwd_bwd_host", "nz_scan_num_chunks", "nz_last_error"} <= set(syms)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/nnuzoo_b200.h but not exported"


def test_struct_mirror_matches_the_compiled_layout():
    from nnuzoo_b200 import _native
    lib = _native.lib()  # raises if sizeof(NzScanDesc) or the ABI version disagree
    assert lib.nz_sizeof_scan_desc() == ctypes.sizeof(_native.NzScanDesc)
    ck = _native.NZ_CHUNK
    assert lib.nz_scan_num_chunks(1) == 1 and lib.nz_scan_num_chunks(ck + 1) == 2
    assert lib.nz_scan_num_chunks(512 * 512) == 512 * 512 // ck
    d = _native.NzScanDesc()
    d.batch, d.dim = 3, 40
    assert lib.nz_scan_workspace_bytes(ctypes.byref(d)) == _native.workspace_bytes(3, 40)


def test_row_per_lane_abi_queries():
    """nz_scan_fine_bytes / nz_scan_workspace_bytes_bwd / nz_scan_bwd_overwrites_dbc are pure host logic."""
    from nnuzoo_b200 import _native
    lib = _native.lib()
    d = _native.NzScanDesc()
    d.batch, d.dim, d.dstate, d.ngroups, d.seqlen, d.dtype = 12, 128, 16, 4, 65536, 0
    d.u = d.delta = d.B = d.C = 256
    for st in (d.u_stride, d.delta_stride):
        st[0], st[1] = 128 * 65536, 65536
    for st in (d.B_stride, d.C_stride):
        st[0], st[1], st[2] = 4 * 16 * 65536, 16 * 65536, 65536
    assert lib.nz_scan_fine_bytes(ctypes.byref(d)) == 12 * 128 * (65536 // 8) * 16 * 4
    assert lib.nz_scan_bwd_overwrites_dbc(ctypes.byref(d)) == 0       # warp-scan backward, 32 rows per group: atomics
    d.xf = 256
    assert lib.nz_scan_bwd_overwrites_dbc(ctypes.byref(d)) == 1       # row-per-lane, one warp per group
    assert lib.nz_scan_workspace_bytes_bwd(ctypes.byref(d)) > lib.nz_scan_workspace_bytes(ctypes.byref(d))
    d.ngroups = 16                                                    # 8 rows per group: not eligible
    d.xf = None
    assert lib.nz_scan_fine_bytes(ctypes.byref(d)) == 0
    assert lib.nz_scan_bwd_overwrites_dbc(ctypes.byref(d)) == 1       # warp-scan backward, <= 16 rows per group


def test_invalid_descriptors_are_rejected_with_a_message():
    from nnuzoo_b200 import _native
    lib = _native.lib()
    d = _native.NzScanDesc()  # all zero: invalid dims; validation happens before any CUDA call
    assert lib.nz_scan_fwd(ctypes.byref(d), None) == -1
    assert b"must be >= 1" in lib.nz_last_error()
    d.batch = d.dim = d.ngroups = 1
    d.seqlen = 8
    d.dstate = 32
    assert lib.nz_scan_fwd(ctypes.byref(d), None) == -2  # NZ_EUNSUPPORTED: d_state > 16
    assert lib.nz_cross_scan(None, None, 0, 1, 1, 2, None, None) == -1


def test_product_path_has_no_cpu_fallback_and_never_touches_the_oracle():
    import nnuzoo_b200
    u = torch.zeros(1, 8, 16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        nnuzoo_b200.selective_scan_fn(u, u, torch.zeros(8, 16), torch.zeros(1, 1, 16, 16), torch.zeros(1, 1, 16, 16))
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        nnuzoo_b200.cross_scan(torch.zeros(1, 2, 4, 4))
    pkg = os.path.join(ROOT, "nnuzoo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
    code = "import sys, nnuzoo_b200; assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules)"
    subprocess.check_call([sys.executable, "-c", code], cwd=ROOT)


def test_missing_library_fails_loudly(tmp_path):
    code = ("import os; os.environ['NNUZOO_B200_LIB'] = r'%s'\n"
            "from nnuzoo_b200 import _native\n"
            "try:\n    _native.lib()\nexcept _native.NativeLibraryError as e:\n    print('LOUD', e)\n") % str(tmp_path / "nope.so")
    out = subprocess.check_output([sys.executable, "-c", code], cwd=ROOT, text=True)
    assert "LOUD" in out and "no CPU or PyTorch fallback" in out


def test_modules_keep_the_reference_state_dict_contract():
    """m2net.py:69-110 / ssnd2net.py:108-179 parameter names and shapes."""
    from nnuzoo_b200 import SS2D, SSND
    m = SS2D(d_model=32)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert shapes == {
        "x_proj_weight": (4, 2 + 32, 64), "dt_projs_weight": (4, 64, 2), "dt_projs_bias": (4, 64),
        "A_logs": (256, 16), "Ds": (256,), "in_proj.weight": (128, 32), "conv2d.weight": (64, 1, 3, 3),
        "conv2d.bias": (64,), "out_norm.weight": (64,), "out_norm.bias": (64,), "out_proj.weight": (32, 64)}
    assert m.A_logs._no_weight_decay and m.Ds._no_weight_decay
    assert torch.allclose(m.A_logs[0].exp(), torch.arange(1, 17).float())
    s = SSND(3, "cross-scan", d_model=16)
    assert s.k == 6 and tuple(s.x_proj_weight.shape) == (6, 1 + 32, 32)
    assert "convnd.conv.weight" in s.state_dict() and tuple(s.state_dict()["convnd.conv.weight"].shape) == (32, 1, 3, 3, 3)
    with pytest.raises(Exception):
        SSND(2, "other", d_model=8)


def test_mamba_block_loads_the_reference_state_dict():
    """seg_mamba/mamba_simple.py:37-189: same parameter names / shapes, incl. the *_b and *_s sets the
    reference creates unconditionally (checked against a state_dict recorded from the reference)."""
    import numpy as np

    from nnuzoo_b200 import Mamba
    rec = np.load(os.path.join(ROOT, "tests", "golden", "module_mamba_v3.npz"))
    sd = {k[3:]: torch.from_numpy(rec[k]) for k in rec.files if k.startswith("sd_")}
    m = Mamba(d_model=32, bimamba_type="v3", nslices=5)
    m.load_state_dict(sd, strict=True)
    assert m.dt_rank == 2 and m.d_inner == 64
    assert m.A_log._no_weight_decay and m.D_s._no_weight_decay and m.dt_proj.bias._no_reinit
    fresh = Mamba(d_model=32)
    assert torch.allclose(fresh.A_b_log[0].exp(), torch.arange(1, 17).float())
    sp = torch.nn.functional.softplus(fresh.dt_proj.bias)
    assert float(sp.min()) >= 1e-4 - 1e-9 and float(sp.max()) <= 0.1 + 1e-6
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 10, 32))


def test_ss2d_accepts_every_reference_constructor_spelling():
    """m2net.py:40-58, SwinUMamba.py:91-113 (ssm_ratio, act_layer, **kwargs), SwinUMambaD.py:155-174 (**kwargs),
    LightSS2DMambaUNet.py:78-96 and the VSSBlock call SS2D(d_model=..., dropout=..., d_state=..., **kwargs)."""
    import torch.nn as nn

    from nnuzoo_b200 import SS2D
    a = SS2D(d_model=32, dropout=0.0, d_state=16)
    b = SS2D(d_model=32, ssm_ratio=2, act_layer=nn.SiLU, d_state=16, some_future_kwarg=1)
    assert {k: tuple(v.shape) for k, v in a.state_dict().items()} == {k: tuple(v.shape) for k, v in b.state_dict().items()}
    c = SS2D(d_model=32, ssm_ratio=1)
    assert c.d_inner == 32 and SS2D().d_model == 96
    with pytest.raises(NotImplementedError):
        SS2D(d_model=32, act_layer=nn.GELU)
