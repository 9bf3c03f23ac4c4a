"""ctypes binding of libnnuzoo_b200.so (the C ABI of include/nnuzoo_b200.h).

There is deliberately no fallback: if the CUDA library is missing or fails to load, importing the
ops raises.  The product path never touches oracle/.
"""
from __future__ import annotations

import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
# NNUZOO_B200_LIB lets the tuning tools load an alternative build of the same ABI
LIB_PATH = os.environ.get("NNUZOO_B200_LIB") or os.path.join(_HERE, "lib", "libnnuzoo_b200.so")

NZ_F32, NZ_BF16, NZ_F16 = 0, 1, 2
NZ_CHUNK = int(os.environ.get("NNUZOO_B200_CHUNK", "128"))   # tuning builds may use another interval
NZ_MAX_DSTATE = 16
ABI_VERSION = 4
NZ_FINE = 8  # steps between two fine checkpoints (NzScanDesc.xf)
WS_HEADER = 256

_vp = ctypes.c_void_p
_i32 = ctypes.c_int32
_i64 = ctypes.c_int64


class NzScanDesc(ctypes.Structure):
    """Field-for-field mirror of ``struct NzScanDesc`` in include/nnuzoo_b200.h."""

    _fields_ = [
        ("batch", _i32), ("dim", _i32), ("dstate", _i32), ("ngroups", _i32),
        ("seqlen", _i64),
        ("dtype", _i32), ("delta_softplus", _i32), ("force_generic", _i32), ("out_f32", _i32),
        ("u", _vp), ("delta", _vp), ("A", _vp), ("B", _vp), ("C", _vp), ("D", _vp), ("z", _vp),
        ("delta_bias", _vp),
        ("u_stride", _i64 * 2), ("delta_stride", _i64 * 2), ("z_stride", _i64 * 2),
        ("B_stride", _i64 * 3), ("C_stride", _i64 * 3), ("A_stride", _i64),
        ("out", _vp), ("out_stride", _i64 * 2), ("x", _vp),
        ("dout", _vp), ("dout_stride", _i64 * 2),
        ("du", _vp), ("ddelta", _vp), ("dz", _vp), ("dA", _vp), ("dB", _vp), ("dC", _vp), ("dD", _vp),
        ("ddelta_bias", _vp),
        ("workspace", _vp), ("workspace_bytes", _i64),
        ("xf", _vp),
        ("rev_mask", _i32), ("u_gdiv", _i32),
    ]


class NzConv1dDesc(ctypes.Structure):
    """Field-for-field mirror of ``struct NzConv1dDesc`` in include/nnuzoo_b200.h."""

    _fields_ = [
        ("batch", _i32), ("dim", _i32), ("width", _i32), ("dtype", _i32),
        ("seqlen", _i64),
        ("silu", _i32), ("reverse", _i32),
        ("x", _vp), ("weight", _vp), ("bias", _vp), ("out", _vp), ("dout", _vp), ("dx", _vp),
        ("dweight", _vp), ("dbias", _vp),
        ("x_stride", _i64 * 2), ("out_stride", _i64 * 2), ("dout_stride", _i64 * 2),
    ]


_lib = None
_lock = threading.Lock()


class NativeLibraryError(RuntimeError):
    pass


def lib():
    """Load (once) and return the native library; raise loudly if it is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NativeLibraryError(
                        f"{LIB_PATH} is missing: build it with `python -m nnuzoo_b200.build` "
                        "(there is no CPU or PyTorch fallback for this path)")
                L = ctypes.CDLL(LIB_PATH)
                L.nz_scan_num_chunks.argtypes = [_i64]
                L.nz_scan_num_chunks.restype = _i64
                L.nz_scan_workspace_bytes.argtypes = [ctypes.POINTER(NzScanDesc)]
                L.nz_scan_workspace_bytes.restype = _i64
                L.nz_scan_workspace_bytes_cp.argtypes = [ctypes.POINTER(NzScanDesc)]
                L.nz_scan_workspace_bytes_cp.restype = _i64
                for name in ("nz_scan_fine_bytes", "nz_scan_workspace_bytes_bwd"):
                    fn = getattr(L, name)
                    fn.argtypes = [ctypes.POINTER(NzScanDesc)]
                    fn.restype = _i64
                L.nz_scan_bwd_overwrites_dbc.argtypes = [ctypes.POINTER(NzScanDesc)]
                L.nz_scan_bwd_overwrites_dbc.restype = ctypes.c_int
                for name in ("nz_scan_fwd", "nz_scan_bwd", "nz_scan_fwd_bwd_host"):
                    fn = getattr(L, name)
                    fn.argtypes = [ctypes.POINTER(NzScanDesc), _vp]
                    fn.restype = ctypes.c_int
                L.nz_cross_scan.argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, ctypes.POINTER(_i64), _vp]
                L.nz_cross_scan.restype = ctypes.c_int
                for name in ("nz_cross_merge", "nz_cross_merge_bwd"):
                    fn = getattr(L, name)
                    fn.argtypes = [_vp, _vp, _i32, _i32, _i32, ctypes.POINTER(_i64), _i32, _vp]
                    fn.restype = ctypes.c_int
                for name in ("nz_causal_conv1d_fwd", "nz_causal_conv1d_bwd"):
                    fn = getattr(L, name)
                    fn.argtypes = [ctypes.POINTER(NzConv1dDesc), _vp]
                    fn.restype = ctypes.c_int
                L.nz_proj_wgrad.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i64,
                                            ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]
                L.nz_proj_wgrad.restype = ctypes.c_int
                L.nz_layernorm_supported.argtypes = [_i32, _i32, _i32]
                L.nz_layernorm_supported.restype = ctypes.c_int
                L.nz_layernorm_fwd.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, ctypes.c_float, _vp]
                L.nz_layernorm_fwd.restype = ctypes.c_int
                L.nz_layernorm_bwd.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp]
                L.nz_layernorm_bwd.restype = ctypes.c_int
                L.nz_dwconv3x3_fwd.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
                L.nz_dwconv3x3_fwd.restype = ctypes.c_int
                L.nz_dwconv3x3_bwd.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
                L.nz_dwconv3x3_bwd.restype = ctypes.c_int
                L.nz_ss2d_epilogue_supported.argtypes = [_i32]
                L.nz_ss2d_epilogue_supported.restype = ctypes.c_int
                L.nz_ss2d_epilogue_fwd.argtypes = [_vp, _vp, ctypes.POINTER(_i64), _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32,
                                                   _i32, _i32, _i32, _i32, ctypes.c_float, _vp]
                L.nz_ss2d_epilogue_fwd.restype = ctypes.c_int
                L.nz_ss2d_epilogue_bwd.argtypes = [_vp, _vp, _vp, _vp, _vp, ctypes.POINTER(_i64), _vp, _vp, _vp, _vp, _vp,
                                                   _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]
                L.nz_ss2d_epilogue_bwd.restype = ctypes.c_int
                for name in ("nz_ss2d_epilogue_fwd", "nz_ss2d_epilogue_bwd"):
                    getattr(L, name + "_folded").argtypes = getattr(L, name).argtypes
                    getattr(L, name + "_folded").restype = ctypes.c_int
                for name in ("nz_cross_scan_pair", "nz_cross_merge_pair"):
                    getattr(L, name).argtypes = [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]
                    getattr(L, name).restype = ctypes.c_int
                L.nz_sw_accumulate.argtypes = [_vp, _i32, _i32, ctypes.POINTER(_i32), _i32, _i32, ctypes.POINTER(_i64), _vp,
                                               _i32, _vp, _vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64), _vp]
                L.nz_sw_accumulate.restype = ctypes.c_int
                L.nz_sizeof_conv1d_desc.restype = _i64
                if L.nz_sizeof_conv1d_desc() != ctypes.sizeof(NzConv1dDesc):
                    raise NativeLibraryError("NzConv1dDesc layout differs between _native.py and the .so")
                L.nz_last_error.restype = ctypes.c_char_p
                L.nz_abi_version.restype = ctypes.c_int
                L.nz_set_device.argtypes = [ctypes.c_int]
                L.nz_set_device.restype = ctypes.c_int
                L.nz_launch_count.restype = _i64
                L.nz_sizeof_scan_desc.restype = _i64
                L.nz_debug_set_trace.argtypes = [_vp]
                L.nz_debug_set_trace.restype = None
                if L.nz_sizeof_scan_desc() != ctypes.sizeof(NzScanDesc):
                    raise NativeLibraryError("NzScanDesc layout differs between _native.py and the .so")
                if L.nz_abi_version() != ABI_VERSION:
                    raise NativeLibraryError("ABI version mismatch between _native.py and the .so")
                _lib = L
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().nz_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def bind_device(index: int) -> None:
    """Make `index` current in the library's own CUDA runtime for the calling thread."""
    # unconditional: another library (PyTorch included) may have switched the thread's current device since the last
    # call, and the call is cheap
    check(lib().nz_set_device(int(index)), "nz_set_device")


def workspace_bytes(batch: int, dim: int) -> int:
    """Scratch bytes of one scan call (mirror of nz_scan_workspace_bytes; checked in tests/test_abi.py)."""
    return WS_HEADER + batch * dim * 2 * NZ_MAX_DSTATE * 8


def launch_count() -> int:
    return int(lib().nz_launch_count())
