"""CrossScan / CrossMerge of SS2D (2-D, 4 directions) and SSND (3-D, 6 directions) as CUDA ops.

The reference writes these as inline PyTorch tensor shuffles that materialise several copies:
  scan  2-D: nnunetv2/nets/m2net.py:175-177        3-D: nnunetv2/nets/ssnd2net.py:250-255
  merge 2-D: nnunetv2/nets/m2net.py:202-206, :218  3-D: nnunetv2/nets/ssnd2net.py:286-298
Here each is one bit-exact kernel (nz_cross_scan / nz_cross_merge in include/nnuzoo_b200.h) with
the matching adjoint for autograd.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native

_DTYPES = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def _spatial(shape):
    arr = (ctypes.c_int64 * len(shape))(*[int(s) for s in shape])
    return arr


def _stream(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _need_cuda(t, name):
    if not t.is_cuda:
        raise RuntimeError(f"nnuzoo_b200.{name}: CUDA tensors only (no CPU fallback)")


def _merge_raw(out_y, spatial, mode):
    b, k, d, L = out_y.shape
    y = torch.empty((b, d, L), dtype=torch.float32, device=out_y.device)
    _native.bind_device(out_y.device.index)
    with torch.cuda.device(out_y.device):
        _native.check(_native.lib().nz_cross_merge(
            ctypes.c_void_p(out_y.data_ptr()), ctypes.c_void_p(y.data_ptr()), b, d, len(spatial),
            _spatial(spatial), mode, _stream(out_y.device)), "nz_cross_merge")
    return y


class CrossScanFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        _need_cuda(x, "cross_scan")
        if x.dtype not in _DTYPES:
            raise TypeError(f"unsupported dtype {x.dtype}")
        x = x.contiguous()
        b, d = x.shape[:2]
        spatial = tuple(x.shape[2:])
        if len(spatial) not in (2, 3):
            raise ValueError("cross_scan expects (B, D, H, W) or (B, D, Z, H, W)")
        L = 1
        for s in spatial:
            L *= s
        xs = torch.empty((b, 2 * len(spatial), d, L), dtype=x.dtype, device=x.device)
        _native.bind_device(x.device.index)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().nz_cross_scan(
                ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(xs.data_ptr()), _DTYPES[x.dtype], b, d,
                len(spatial), _spatial(spatial), _stream(x.device)), "nz_cross_scan")
        ctx.spatial = spatial
        ctx.in_dtype = x.dtype
        return xs

    @staticmethod
    def backward(ctx, dxs):
        # adjoint of a gather = sum of the K un-permuted gradients = the "all directions" merge
        dx = _merge_raw(dxs.float().contiguous(), ctx.spatial, 1)
        b, d = dx.shape[:2]
        return dx.view(b, d, *ctx.spatial).to(ctx.in_dtype)


class CrossMergeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, out_y, spatial, mode):
        _need_cuda(out_y, "cross_merge")
        if out_y.dtype != torch.float32:
            raise TypeError("cross_merge sums in fp32 like the reference (m2net.py:200 asserts float)")
        out_y = out_y.contiguous()
        if out_y.shape[1] != 2 * len(spatial):
            raise ValueError("out_y must be (B, 2 * len(spatial), D, L)")
        ctx.spatial, ctx.mode = tuple(spatial), int(mode)
        return _merge_raw(out_y, ctx.spatial, ctx.mode)

    @staticmethod
    def backward(ctx, dy):
        dy = dy.float().contiguous()
        b, d, L = dy.shape
        k = 2 * len(ctx.spatial)
        doy = torch.empty((b, k, d, L), dtype=torch.float32, device=dy.device)
        _native.bind_device(dy.device.index)
        with torch.cuda.device(dy.device):
            _native.check(_native.lib().nz_cross_merge_bwd(
                ctypes.c_void_p(dy.data_ptr()), ctypes.c_void_p(doy.data_ptr()), b, d, len(ctx.spatial),
                _spatial(ctx.spatial), ctx.mode, _stream(dy.device)), "nz_cross_merge_bwd")
        return doy, None, None


def cross_scan(x: torch.Tensor) -> torch.Tensor:
    """x (B, D, H, W) -> (B, 4, D, L)  |  x (B, D, Z, H, W) -> (B, 6, D, L); bit-exact vs the reference."""
    return CrossScanFn.apply(x)


def cross_merge(out_y: torch.Tensor, spatial, mode: str = "reference") -> torch.Tensor:
    """out_y (B, K, D, L) fp32 -> y (B, D, L) fp32 in row-major spatial order, reference sum order.

    mode="reference" reproduces ssnd2net.py:291-298 bit-exactly in 3-D (directions 2 and 5 unused);
    mode="fixed" un-permutes direction 2 / 5 from their own order (never used for parity)."""
    return CrossMergeFn.apply(out_y, tuple(int(s) for s in spatial), {"reference": 0, "fixed": 1}[mode])


class CrossScanPairFn(torch.autograd.Function):
    """x (B, D, H, W) -> xs2 (B, 2, D, L) = {row-major walk, column-major walk}: the two arrays SS2D's four directions
    walk (m2net.py:175-177); the L-flipped directions are walked backwards by the scan itself (``rev_mask``) instead of
    being materialised.  Backward = nz_cross_merge_pair."""

    @staticmethod
    def forward(ctx, x):
        _need_cuda(x, "cross_scan_pair")
        if x.dtype not in _DTYPES or x.dim() != 4:
            raise TypeError("cross_scan_pair expects a (B, D, H, W) fp32 / bf16 / fp16 tensor")
        x = x.contiguous()
        b, d, H, W = x.shape
        xs2 = torch.empty((b, 2, d, H * W), dtype=x.dtype, device=x.device)
        _native.bind_device(x.device.index)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().nz_cross_scan_pair(
                ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(xs2.data_ptr()), _DTYPES[x.dtype], b, d, H, W,
                _stream(x.device)), "nz_cross_scan_pair")
        ctx.hw = (H, W)
        return xs2

    @staticmethod
    def backward(ctx, dxs2):
        H, W = ctx.hw
        dxs2 = dxs2.contiguous()
        b, _, d, _ = dxs2.shape
        dx = torch.empty((b, d, H, W), dtype=dxs2.dtype, device=dxs2.device)
        _native.bind_device(dxs2.device.index)
        with torch.cuda.device(dxs2.device):
            _native.check(_native.lib().nz_cross_merge_pair(
                ctypes.c_void_p(dxs2.data_ptr()), ctypes.c_void_p(dx.data_ptr()), _DTYPES[dxs2.dtype], b, d, H, W,
                _stream(dxs2.device)), "nz_cross_merge_pair")
        return dx


def cross_scan_pair(x: torch.Tensor) -> torch.Tensor:
    """x (B, D, H, W) -> (B, 2, D, L): row-major and column-major walks only (see CrossScanPairFn)."""
    return CrossScanPairFn.apply(x)
