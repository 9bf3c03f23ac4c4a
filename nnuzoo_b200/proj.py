"""Per-direction 1x1 projections of SS2D with a hand-written weight-gradient kernel.

``grouped_proj(x, w)`` is the reference's ``torch.einsum("b k n l, k m n -> b k m l", x, w)`` -- the x_proj at
nnunetv2/nets/m2net.py:179 (n = d_inner, m = R + 2N) and the dt_proj at :182 (n = R, m = d_inner).  Forward and
input gradient are plain batched GEMMs (cuBLAS: L is the long, parallel axis).  The weight gradient is a reduction
over batch x L = millions of positions into a 33 x 32 matrix, the shape library GEMMs serve worst (one tile per
direction, 48 % of an M2Net training step before this op existed); it runs on ``nz_proj_wgrad``
(csrc/proj_kernels.cu) through the C ABI.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native

_DT = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def proj_wgrad(g: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """dW (K, M, N) fp32 = sum over (b, l) of g (B, K, M, L) x (B, K, N, L)."""
    if not (g.is_cuda and x.is_cuda):
        raise RuntimeError("nnuzoo_b200.proj_wgrad: CUDA tensors required (this path has no CPU fallback)")
    if g.dim() != 4 or x.dim() != 4 or g.shape[:2] != x.shape[:2] or g.shape[3] != x.shape[3]:
        raise ValueError(f"proj_wgrad: incompatible shapes {tuple(g.shape)} / {tuple(x.shape)}")
    if g.dtype not in _DT or x.dtype not in _DT:
        raise TypeError(f"proj_wgrad: unsupported dtypes {g.dtype} / {x.dtype}")
    if g.stride(3) != 1:
        g = g.contiguous()
    if x.stride(3) != 1:
        x = x.contiguous()
    B, K, M, L = g.shape
    N = x.shape[2]
    dw = torch.zeros(K, M, N, dtype=torch.float32, device=g.device)
    gs = (ctypes.c_int64 * 3)(*g.stride()[:3])
    xs = (ctypes.c_int64 * 3)(*x.stride()[:3])
    _native.bind_device(g.device.index)
    stream = ctypes.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
    rc = _native.lib().nz_proj_wgrad(ctypes.c_void_p(g.data_ptr()), ctypes.c_void_p(x.data_ptr()),
                                     ctypes.c_void_p(dw.data_ptr()), _DT[g.dtype], _DT[x.dtype], B, K, M, N, L, gs, xs,
                                     stream)
    if rc == -2:
        # NZ_EUNSUPPORTED: M * N beyond what one CTA's accumulators hold (wider SS2D than M2Net's, e.g. SwinUMamba's
        # d_model 192: 44 x 384).  The projection is GEMM-shaped, so the library GEMM is the plain path here -- as
        # norm.py does for channel counts its kernel does not cover.
        return torch.einsum("bkml,bknl->kmn", g.float(), x.float())
    _native.check(rc, "nz_proj_wgrad")
    return dw


class GroupedProjFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w):
        ctx.save_for_backward(x, w)
        return torch.matmul(w.to(x.dtype).unsqueeze(0), x)            # (1, K, M, N) @ (B, K, N, L)

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        dx = dw = None
        if ctx.needs_input_grad[0]:
            dx = torch.matmul(w.to(g.dtype).transpose(1, 2).unsqueeze(0), g).to(x.dtype)
        if ctx.needs_input_grad[1]:
            dw = proj_wgrad(g, x).to(w.dtype)
        return dx, dw


def grouped_proj(x: torch.Tensor, w: torch.Tensor) -> torch.Tensor:
    """x (B, K, N, L), w (K, M, N) -> (B, K, M, L) in x's dtype; the reference's einsum at m2net.py:179 / :182."""
    return GroupedProjFn.apply(x, w)
