"""ctypes front-end for oracle/scan_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  It restates (on the CPU, in C) the arithmetic of the
reference's ``selective_scan_ref``
(nnunetv2/nets/seg_mamba/selective_scan_interface.py:86-152) and its adjoint; the
numpy-facing signature mirrors ``selective_scan_ref`` argument for argument.

Parity status: pinned against the reference's own outputs (tests/golden/, made by
oracle/gen_golden.py in the build container where /root/reference is mounted).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libscan_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle with the committed Makefile (gcc only, no GPU needed)."""
    if force or not os.path.exists(_LIB_PATH) or (
        os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "scan_oracle.c"))
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B" if force else "all"])
    return _LIB_PATH


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_LIB_PATH)
        fp = ctypes.POINTER(ctypes.c_float)
        fwd_args = [fp, fp, fp, fp, fp, fp, fp, fp, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                    ctypes.c_int, ctypes.c_long, ctypes.c_int, fp, fp]
        bwd_args = [fp, fp, fp, fp, fp, fp, fp, fp, ctypes.c_int, fp, ctypes.c_int, ctypes.c_int,
                    ctypes.c_int, ctypes.c_long, ctypes.c_int, fp, fp, fp, fp, fp, fp, fp, fp]
        for sfx in ("f32", "f64"):
            getattr(_lib, f"nzo_scan_fwd_{sfx}").argtypes = fwd_args
            getattr(_lib, f"nzo_scan_fwd_{sfx}").restype = ctypes.c_int
            getattr(_lib, f"nzo_scan_bwd_{sfx}").argtypes = bwd_args
            getattr(_lib, f"nzo_scan_bwd_{sfx}").restype = ctypes.c_int
        _lib.nzo_num_threads.restype = ctypes.c_int
        _lib.nzo_set_num_threads.argtypes = [ctypes.c_int]
    return _lib


def num_threads() -> int:
    return int(_load().nzo_num_threads())


def set_num_threads(n: int) -> None:
    _load().nzo_set_num_threads(int(n))


def _f32(a) -> Optional[np.ndarray]:
    if a is None:
        return None
    if hasattr(a, "detach"):  # torch tensor (any float dtype) -> fp32 numpy, as :102-118 do
        a = a.detach().float().cpu().numpy()
    return np.ascontiguousarray(a, dtype=np.float32)


def _ptr(a: Optional[np.ndarray]):
    if a is None:
        return None
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _canon(u, delta, A, B, C, D, z, delta_bias):
    u, delta, A, B, C, D, z, delta_bias = (_f32(t) for t in (u, delta, A, B, C, D, z, delta_bias))
    batch, dim, L = u.shape
    dstate = A.shape[1]
    squeeze_b = B.ndim == 3
    squeeze_c = C.ndim == 3
    if squeeze_b:
        B = B[:, None]
    if squeeze_c:
        C = C[:, None]
    if B.shape[1] != C.shape[1]:
        raise ValueError("oracle: B and C must have the same number of groups")
    ngroups = B.shape[1]
    assert A.shape[0] == dim and B.shape == (batch, ngroups, dstate, L) and C.shape == B.shape
    return (u, delta, A, np.ascontiguousarray(B), np.ascontiguousarray(C), D, z, delta_bias,
            batch, dim, dstate, L, ngroups, squeeze_b, squeeze_c)


def selective_scan_oracle(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                          return_last_state=False, precision: str = "f32"):
    """Forward; same arguments/returns as selective_scan_ref (:86-152), numpy fp32 results."""
    (u, delta, A, B, C, D, z, delta_bias, batch, dim, dstate, L, ngroups, _, _) = _canon(
        u, delta, A, B, C, D, z, delta_bias)
    out = np.empty((batch, dim, L), np.float32)
    last = np.empty((batch, dim, dstate), np.float32)
    fn = getattr(_load(), f"nzo_scan_fwd_{precision}")
    rc = fn(_ptr(u), _ptr(delta), _ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(z), _ptr(delta_bias),
            int(bool(delta_softplus)), batch, dim, dstate, L, ngroups, _ptr(out), _ptr(last))
    if rc:
        raise RuntimeError(f"nzo_scan_fwd_{precision} failed rc={rc}")
    return (out, last) if return_last_state else out


def selective_scan_oracle_bwd(u, delta, A, B, C, D, z, delta_bias, delta_softplus, dout,
                              precision: str = "f32"):
    """Adjoint. Returns dict(du, ddelta, dA, dB, dC, dD, dz, ddelta_bias) (None where the input
    was None), shaped like the corresponding inputs -- the tuple the reference's
    SelectiveScanFn.backward returns at selective_scan_interface.py:69-74."""
    (u, delta, A, B, C, D, z, delta_bias, batch, dim, dstate, L, ngroups, sq_b, sq_c) = _canon(
        u, delta, A, B, C, D, z, delta_bias)
    dout = _f32(dout)
    du = np.empty_like(u)
    ddelta = np.empty_like(u)
    dA = np.empty_like(A)
    dB = np.empty_like(B)
    dC = np.empty_like(C)
    dD = np.empty((dim,), np.float32)
    dbias = np.empty((dim,), np.float32)
    dz = np.empty_like(u) if z is not None else None
    fn = getattr(_load(), f"nzo_scan_bwd_{precision}")
    rc = fn(_ptr(u), _ptr(delta), _ptr(A), _ptr(B), _ptr(C), _ptr(D), _ptr(z), _ptr(delta_bias),
            int(bool(delta_softplus)), _ptr(dout), batch, dim, dstate, L, ngroups,
            _ptr(du), _ptr(ddelta), _ptr(dA), _ptr(dB), _ptr(dC), _ptr(dD), _ptr(dz), _ptr(dbias))
    if rc:
        raise RuntimeError(f"nzo_scan_bwd_{precision} failed rc={rc}")
    return dict(du=du, ddelta=ddelta, dA=dA,
                dB=dB[:, 0] if sq_b else dB, dC=dC[:, 0] if sq_c else dC,
                dD=dD if D is not None else None, dz=dz,
                ddelta_bias=dbias if delta_bias is not None else None)
