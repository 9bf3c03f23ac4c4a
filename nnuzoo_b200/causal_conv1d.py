"""Depthwise causal conv1d (+ SiLU) of the 1-D Mamba block as a CUDA op.

Mirrors ``causal_conv1d_fn(x, weight, bias=None, activation=None)`` of the un-vendored
``causal_conv1d`` package the reference calls (mamba_simple.py:319-324; fused form at
selective_scan_interface.py:177, :247-252); semantics are the reference's own fallback
``self.act(self.conv1d(x)[..., :seqlen])`` (mamba_simple.py:316-317).  x: (batch, dim, L) with unit
innermost stride (a chunk view of xz is fine), weight: (dim, width <= 4), bias: (dim) or None.
Runs nz_causal_conv1d_fwd / _bwd (include/nnuzoo_b200.h); no CPU fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native
from ._native import NzConv1dDesc

_DTYPES = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _desc(x, weight, bias, silu, reverse=False):
    d = NzConv1dDesc()
    d.batch, d.dim, d.seqlen = x.shape
    d.width, d.dtype, d.silu, d.reverse = weight.shape[1], _DTYPES[x.dtype], int(silu), int(bool(reverse))
    d.x, d.weight, d.bias = _ptr(x), _ptr(weight), _ptr(bias)
    d.x_stride[0], d.x_stride[1] = x.stride(0), x.stride(1)
    return d


class CausalConv1dFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu, reverse=False):
        if not x.is_cuda:
            raise RuntimeError("nnuzoo_b200.causal_conv1d_fn: CUDA tensors only (no CPU fallback)")
        if x.dtype not in _DTYPES:
            raise TypeError(f"unsupported dtype {x.dtype}")
        if x.dim() != 3 or weight.dim() != 2 or weight.shape[0] != x.shape[1]:
            raise ValueError("x must be (batch, dim, L) and weight (dim, width)")
        if weight.shape[1] > 4:
            raise NotImplementedError("causal_conv1d width > 4 is not implemented (nnUZoo uses d_conv = 4)")
        ctx.in_dtypes = (weight.dtype, None if bias is None else bias.dtype)
        if x.stride(-1) != 1:
            x = x.contiguous()
        w32 = weight.float().contiguous()
        b32 = None if bias is None else bias.float().contiguous()
        out = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        d = _desc(x, w32, b32, silu, reverse)
        d.out = _ptr(out)
        d.out_stride[0], d.out_stride[1] = out.stride(0), out.stride(1)
        _native.bind_device(x.device.index)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().nz_causal_conv1d_fwd(
                ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "nz_causal_conv1d_fwd")
        ctx.save_for_backward(x, w32, b32)
        ctx.silu, ctx.reverse = bool(silu), bool(reverse)
        return out

    @staticmethod
    def backward(ctx, dout):
        x, w32, b32 = ctx.saved_tensors
        if dout.stride(-1) != 1 or dout.dtype != x.dtype:
            dout = dout.to(x.dtype).contiguous()
        dx = torch.empty(x.shape, dtype=x.dtype, device=x.device)
        dw = torch.zeros_like(w32)
        db = None if b32 is None else torch.zeros_like(b32)
        d = _desc(x, w32, b32, ctx.silu, ctx.reverse)
        d.dout, d.dx, d.dweight, d.dbias = _ptr(dout), _ptr(dx), _ptr(dw), _ptr(db)
        d.dout_stride[0], d.dout_stride[1] = dout.stride(0), dout.stride(1)
        _native.bind_device(x.device.index)
        with torch.cuda.device(x.device):
            _native.check(_native.lib().nz_causal_conv1d_bwd(
                ctypes.byref(d), ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)), "nz_causal_conv1d_bwd")
        wt, bt = ctx.in_dtypes
        return dx, dw.to(wt), (None if db is None else db.to(bt)), None, None


def causal_conv1d_fn(x, weight, bias=None, activation=None, *, reverse=False):
    """x: (batch, dim, seqlen); weight: (dim, width); bias: (dim,); activation: None | "silu" | "swish".

    Extension (keyword-only, not in the reference signature): ``reverse=True`` computes
    ``causal_conv1d_fn(x.flip(-1), ...).flip(-1)`` -- the convolution of the reversed Mamba directions
    (mamba_simple.py:250-262, mamba_nd2net.py:638-641) -- without making either flipped copy."""
    if activation not in (None, "silu", "swish"):
        raise NotImplementedError("activation must be None, silu or swish")
    return CausalConv1dFn.apply(x, weight, bias, activation is not None, reverse)
