"""Channels-last LayerNorm on the sm_100a kernel (csrc/norm_kernels.cu).

``LayerNorm`` is a drop-in for the ``nn.LayerNorm`` instances of the SS2D nets (same parameters / state-dict keys):
ln_1 and out_norm of every VSS block (nnunetv2/nets/m2net.py:524, :220), patch merge / expand norms (:241, :286-290).
Dtype rule = the reference's: statistics and affine in fp32; without autocast the output has the input's dtype; under
``torch.autocast`` nn.LayerNorm returns fp32 whatever comes in, and so does this one -- except that a norm marked
``feeds_linear`` (ln_1 -> in_proj, PatchMerging2D.norm -> reduction) writes the autocast dtype directly: the Linear
behind it would round the fp32 result to that dtype anyway, so the numbers are bit-identical and one fp32 round trip
through HBM disappears.  Input and output element types are independent in the kernel (bf16 -> fp32, fp32 -> bf16, ...).
Shapes the kernel does not cover (C / E not a power of two) and CPU tensors go to ``F.layer_norm`` -- a library op with
the same semantics, outside the scan path.
"""
from __future__ import annotations

import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native

_DT = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps, out_dtype):
        shape = x.shape
        C = shape[-1]
        x2 = x.contiguous().view(-1, C)
        rows = x2.shape[0]
        y = torch.empty(x2.shape, dtype=out_dtype, device=x.device)
        mean = torch.empty(rows, dtype=torch.float32, device=x.device)
        rstd = torch.empty(rows, dtype=torch.float32, device=x.device)
        w = weight.float().contiguous() if weight is not None else None
        b = bias.float().contiguous() if bias is not None else None
        _native.bind_device(x.device.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _native.check(_native.lib().nz_layernorm_fwd(_vp(x2), _vp(w), _vp(b), _vp(y), _vp(mean), _vp(rstd), rows, C,
                                                     _DT[x.dtype], _DT[out_dtype], float(eps), st), "nz_layernorm_fwd")
        ctx.save_for_backward(x2, w, mean, rstd)
        ctx.out_dtype = out_dtype
        ctx.meta = (shape, weight is not None, bias is not None,
                    weight.dtype if weight is not None else None, bias.dtype if bias is not None else None)
        return y.view(shape)

    @staticmethod
    def backward(ctx, dy):
        x2, w, mean, rstd = ctx.saved_tensors
        shape, has_w, has_b, wdt, bdt = ctx.meta
        rows, C = x2.shape
        dy2 = dy.contiguous().view(rows, C)
        if dy2.dtype != ctx.out_dtype:
            dy2 = dy2.to(ctx.out_dtype)
        dx = torch.empty_like(x2)
        dg = torch.zeros(C, dtype=torch.float32, device=x2.device) if has_w else None
        db = torch.zeros(C, dtype=torch.float32, device=x2.device) if has_b else None
        _native.bind_device(x2.device.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(x2.device).cuda_stream)
        _native.check(_native.lib().nz_layernorm_bwd(_vp(dy2), _vp(x2), _vp(mean), _vp(rstd), _vp(w), _vp(dx), _vp(dg),
                                                     _vp(db), rows, C, _DT[x2.dtype], _DT[ctx.out_dtype], st),
                      "nz_layernorm_bwd")
        return dx.view(shape), (dg.to(wdt) if has_w else None), (db.to(bdt) if has_b else None), None, None


def layer_norm(x: torch.Tensor, weight, bias, eps: float = 1e-5, feeds_linear: bool = False) -> torch.Tensor:
    C = x.shape[-1]
    if x.is_cuda and x.dtype in _DT and x.numel() > 0:
        out_dtype = x.dtype
        if torch.is_autocast_enabled("cuda"):
            out_dtype = torch.get_autocast_dtype("cuda") if feeds_linear else torch.float32
        if out_dtype in _DT and _native.lib().nz_layernorm_supported(C, _DT[x.dtype], _DT[out_dtype]):
            return LayerNormFn.apply(x, weight, bias, eps, out_dtype)
    return F.layer_norm(x, (C,), weight, bias, eps)


class LayerNorm(nn.LayerNorm):
    """nn.LayerNorm over the last axis, executed by nz_layernorm_fwd / _bwd on CUDA tensors.  Set ``feeds_linear``
    on instances whose only consumer is an nn.Linear (see the module docstring)."""

    feeds_linear = False

    def forward(self, x):
        if len(self.normalized_shape) != 1:
            return super().forward(x)
        return layer_norm(x, self.weight, self.bias, self.eps, self.feeds_linear)
