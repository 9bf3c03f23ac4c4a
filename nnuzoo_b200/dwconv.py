"""SS2D's depthwise 3x3 convolution + SiLU as one kernel each way (csrc/dwconv_kernels.cu).

``dwconv3x3_silu(x, weight, bias)`` is the reference's ``self.act(self.conv2d(x))`` (nnunetv2/nets/m2net.py:214-215) for
the SS2D configuration every nnUZoo net uses: kernel 3, padding 1, stride 1, groups = channels, NCHW.  fp32
accumulation; the SiLU is applied to the fp32 pre-activation (the reference rounds the convolution to the autocast
dtype first), output in x's dtype.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native

_DT = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class DwConv3x3SiLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, silu):
        if not x.is_cuda:
            raise RuntimeError("nnuzoo_b200.dwconv3x3_silu: CUDA tensors required (this path has no CPU fallback)")
        if x.dim() != 4 or weight.shape != (x.shape[1], 1, 3, 3) or x.dtype not in _DT:
            raise ValueError(f"dwconv3x3_silu: unsupported shapes / dtype {tuple(x.shape)} {tuple(weight.shape)} {x.dtype}")
        x = x.contiguous()
        w = weight.float().contiguous()
        b = bias.float().contiguous() if bias is not None else None
        y = torch.empty_like(x)
        B, D, H, W = x.shape
        _native.bind_device(x.device.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _native.check(_native.lib().nz_dwconv3x3_fwd(_vp(x), _vp(w), _vp(b), _vp(y), _DT[x.dtype], B, D, H, W,
                                                     int(silu), st), "nz_dwconv3x3_fwd")
        ctx.save_for_backward(x, w, b)
        ctx.silu = int(silu)
        ctx.param_dtypes = (weight.dtype, bias.dtype if bias is not None else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, w, b = ctx.saved_tensors
        dy = dy.contiguous()
        if dy.dtype != x.dtype:
            dy = dy.to(x.dtype)
        B, D, H, W = x.shape
        dx = torch.empty_like(x)
        dw = torch.zeros(D, 1, 3, 3, dtype=torch.float32, device=x.device)
        db = torch.zeros(D, dtype=torch.float32, device=x.device) if b is not None else None
        _native.bind_device(x.device.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _native.check(_native.lib().nz_dwconv3x3_bwd(_vp(x), _vp(dy), _vp(w), _vp(b), _vp(dx), _vp(dw), _vp(db),
                                                     _DT[x.dtype], B, D, H, W, ctx.silu, st), "nz_dwconv3x3_bwd")
        wdt, bdt = ctx.param_dtypes
        return dx, dw.to(wdt), (db.to(bdt) if db is not None else None), None


def dwconv3x3_silu(x: torch.Tensor, weight: torch.Tensor, bias, silu: bool = True) -> torch.Tensor:
    """x (B, D, H, W), weight (D, 1, 3, 3), bias (D) or None -> SiLU(depthwise conv) in x's dtype."""
    return DwConv3x3SiLUFn.apply(x, weight, bias, silu)
