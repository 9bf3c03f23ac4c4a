"""SS2D core as ONE autograd node: selective scan -> CrossMerge -> LayerNorm -> * SiLU(z)   (2-D, 4 directions).

Reference statements: nnunetv2/nets/m2net.py:193-206 (scan + merge), :218-221 (sum, transpose, out_norm, gate).
The node calls ``nz_scan_fwd`` and then ``nz_ss2d_epilogue_fwd`` (csrc/epilogue_kernels.cu); its backward calls
``nz_ss2d_epilogue_bwd`` -- which hands the four permuted copies of dy to the scan in the scan's own operand dtype -- and
then ``nz_scan_bwd``.  Being one node is what allows that hand-over: between separate autograd nodes the engine would cast
the 16-bit gradient back to out_y's fp32.  CUDA only, no fallback.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native
from .selective_scan_interface import SelectiveScanFn

_DT = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def _vp(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else ctypes.c_void_p(0)


class _InnerCtx:
    """Stands in for the autograd ctx of SelectiveScanFn, whose forward / backward bodies are reused as they are."""

    saved_tensors = ()

    def __init__(self, needs_grad: bool = True):
        # SelectiveScanFn.forward asks for fine checkpoints only when some gradient will be wanted
        self.needs_input_grad = (needs_grad,) * 11

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors

    def mark_non_differentiable(self, *tensors):
        pass


def supported(d_inner: int) -> bool:
    return bool(_native.lib().nz_ss2d_epilogue_supported(int(d_inner)))


class SS2DCoreFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, xs, dts, As, Bs, Cs, Ds, dt_bias, z, gamma, beta, H, W, eps, out_dtype):
        bsz, K, D, L = xs.shape
        if K != 4 or H * W != L:
            raise ValueError("ss2d_core: xs must be (B, 4, D, H*W)")
        if z.shape != (bsz, H, W, D) or z.stride(3) != 1 or z.stride(1) != W * z.stride(2) or z.dtype not in _DT:
            raise ValueError("ss2d_core: z must be a (B, H, W, D) tensor with unit channel stride and collapsible H, W")
        inner = _InnerCtx(any(ctx.needs_input_grad[:8]))
        scan_out = torch.float32 if xs.dtype != torch.float32 else None
        out_y = SelectiveScanFn.forward(inner, xs.view(bsz, K * D, L), dts.view(bsz, K * D, L), As, Bs, Cs, Ds, None,
                                        dt_bias, True, False, scan_out)
        dev = xs.device
        out = torch.empty((bsz, H, W, D), dtype=out_dtype, device=dev)
        ym = torch.empty((bsz, L, D), dtype=torch.float32, device=dev)
        mean = torch.empty(bsz * L, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        g32 = gamma.float().contiguous() if gamma is not None else None
        b32 = beta.float().contiguous() if beta is not None else None
        zs = (ctypes.c_int64 * 2)(z.stride(0), z.stride(2))
        _native.bind_device(dev.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _native.check(_native.lib().nz_ss2d_epilogue_fwd(_vp(out_y), _vp(z), zs, _vp(g32), _vp(b32), _vp(out), _vp(ym),
                                                         _vp(mean), _vp(rstd), _DT[z.dtype], _DT[out_dtype], bsz, D, H, W,
                                                         float(eps), st), "nz_ss2d_epilogue_fwd")
        ctx.inner = inner
        ctx.save_for_backward(ym, mean, rstd, z, g32, b32)
        ctx.dims = (bsz, K, D, H, W)
        ctx.dtypes = (xs.dtype, out_dtype, gamma.dtype if gamma is not None else None,
                      beta.dtype if beta is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        ym, mean, rstd, z, g32, b32 = ctx.saved_tensors
        bsz, K, D, H, W = ctx.dims
        op_dtype, out_dtype, gdt, bdt = ctx.dtypes
        L = H * W
        dev = ym.device
        dout = dout.contiguous()
        if dout.dtype != out_dtype:
            dout = dout.to(out_dtype)
        d_out_y = torch.empty((bsz, K * D, L), dtype=op_dtype, device=dev)   # the scan's operand dtype
        dz = torch.empty((bsz, H, W, D), dtype=z.dtype, device=dev)
        dgb = torch.zeros(2 * D, dtype=torch.float32, device=dev)
        dg = dgb[:D] if g32 is not None else None
        db = dgb[D:] if b32 is not None else None
        zs = (ctypes.c_int64 * 2)(z.stride(0), z.stride(2))
        _native.bind_device(dev.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _native.check(_native.lib().nz_ss2d_epilogue_bwd(_vp(dout), _vp(ym), _vp(mean), _vp(rstd), _vp(z), zs, _vp(g32),
                                                         _vp(b32), _vp(d_out_y), _vp(dz), _vp(dg), _vp(db), _DT[z.dtype],
                                                         _DT[out_dtype], _DT[op_dtype], bsz, D, H, W, st),
                      "nz_ss2d_epilogue_bwd")
        du, ddelta, dA, dB, dC, dD, _dz, dbias, *_ = SelectiveScanFn.backward(ctx.inner, d_out_y)
        ctx.inner = None
        return (du.view(bsz, K, D, L), ddelta.view(bsz, K, D, L), dA, dB, dC, dD, dbias, dz,
                dg.to(gdt) if dg is not None else None, db.to(bdt) if db is not None else None, None, None, None, None)


# direction order of the folded layout: k' = 2 * (array: 0 row-major, 1 column-major) + (0 forwards, 1 backwards); in the
# reference's numbering (m2net.py:175-177: k0 row-major, k1 column-major, k2 / k3 their flips) that is k = (0, 2, 1, 3)
FOLD_PERM = (0, 2, 1, 3)
_FOLD_REV_MASK = 0b1010
_FOLD_U_GDIV = 2


class SS2DFoldedFn(torch.autograd.Function):
    """SS2DCoreFn without the flipped copies: the scan reads xs2 (B, 2, D, L) = {x row-major, x column-major}, directions
    k' = 1, 3 walk them from t = L - 1 down to 0 (``rev_mask``), out_y' / d_out_y' (B, 4, D, L) stay un-flipped and the
    epilogue kernels merge / scatter them without index reversal.  dts / Bs / Cs / As / Ds / dt_bias arrive in the folded
    direction order (FOLD_PERM), computed from the un-flipped xs2."""

    @staticmethod
    def forward(ctx, xs2, dts, As, Bs, Cs, Ds, dt_bias, z, gamma, beta, H, W, eps, out_dtype):
        bsz, two, D, L = xs2.shape
        K = 4
        if two != 2 or H * W != L or dts.shape != (bsz, K, D, L):
            raise ValueError("ss2d_core_folded: xs2 must be (B, 2, D, H*W) and dts (B, 4, D, H*W)")
        if z.shape != (bsz, H, W, D) or z.stride(3) != 1 or z.stride(1) != W * z.stride(2) or z.dtype not in _DT:
            raise ValueError("ss2d_core_folded: z must be a (B, H, W, D) tensor with unit channel stride and collapsible H, W")
        inner = _InnerCtx(any(ctx.needs_input_grad[:8]))
        inner.needs_input_grad = inner.needs_input_grad + (False, False)
        scan_out = torch.float32 if xs2.dtype != torch.float32 else None
        out_y = SelectiveScanFn.forward(inner, xs2.view(bsz, 2 * D, L), dts.view(bsz, K * D, L), As, Bs, Cs, Ds, None,
                                        dt_bias, True, False, scan_out, _FOLD_REV_MASK, _FOLD_U_GDIV)
        dev = xs2.device
        out = torch.empty((bsz, H, W, D), dtype=out_dtype, device=dev)
        ym = torch.empty((bsz, L, D), dtype=torch.float32, device=dev)
        mean = torch.empty(bsz * L, dtype=torch.float32, device=dev)
        rstd = torch.empty_like(mean)
        g32 = gamma.float().contiguous() if gamma is not None else None
        b32 = beta.float().contiguous() if beta is not None else None
        zs = (ctypes.c_int64 * 2)(z.stride(0), z.stride(2))
        _native.bind_device(dev.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _native.check(_native.lib().nz_ss2d_epilogue_fwd_folded(
            _vp(out_y), _vp(z), zs, _vp(g32), _vp(b32), _vp(out), _vp(ym), _vp(mean), _vp(rstd), _DT[z.dtype],
            _DT[out_dtype], bsz, D, H, W, float(eps), st), "nz_ss2d_epilogue_fwd_folded")
        ctx.inner = inner
        ctx.save_for_backward(ym, mean, rstd, z, g32, b32)
        ctx.dims = (bsz, K, D, H, W)
        ctx.dtypes = (xs2.dtype, out_dtype, gamma.dtype if gamma is not None else None,
                      beta.dtype if beta is not None else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        ym, mean, rstd, z, g32, b32 = ctx.saved_tensors
        bsz, K, D, H, W = ctx.dims
        op_dtype, out_dtype, gdt, bdt = ctx.dtypes
        L = H * W
        dev = ym.device
        dout = dout.contiguous()
        if dout.dtype != out_dtype:
            dout = dout.to(out_dtype)
        d_out_y = torch.empty((bsz, K * D, L), dtype=op_dtype, device=dev)   # the scan's operand dtype, folded layout
        dz = torch.empty((bsz, H, W, D), dtype=z.dtype, device=dev)
        dgb = torch.zeros(2 * D, dtype=torch.float32, device=dev)
        dg = dgb[:D] if g32 is not None else None
        db = dgb[D:] if b32 is not None else None
        zs = (ctypes.c_int64 * 2)(z.stride(0), z.stride(2))
        _native.bind_device(dev.index)
        st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _native.check(_native.lib().nz_ss2d_epilogue_bwd_folded(
            _vp(dout), _vp(ym), _vp(mean), _vp(rstd), _vp(z), zs, _vp(g32), _vp(b32), _vp(d_out_y), _vp(dz), _vp(dg),
            _vp(db), _DT[z.dtype], _DT[out_dtype], _DT[op_dtype], bsz, D, H, W, st), "nz_ss2d_epilogue_bwd_folded")
        du, ddelta, dA, dB, dC, dD, _dz, dbias, *_ = SelectiveScanFn.backward(ctx.inner, d_out_y)
        ctx.inner = None
        # (du is already summed over the forward and the backward walker of each array: SelectiveScanFn.backward)
        return (du.view(bsz, 2, D, L), ddelta.view(bsz, K, D, L), dA, dB, dC, dD, dbias, dz,
                dg.to(gdt) if dg is not None else None, db.to(bdt) if db is not None else None, None, None, None, None)


class _CoreCtx:
    """Stands in for the autograd ctx of SS2DFoldedFn inside SS2DFoldedProjFn."""

    saved_tensors = ()
    inner = None

    def __init__(self, needs_grad: bool):
        self.needs_input_grad = (needs_grad,) * 8 + (True, True, False, False, False, False)

    def save_for_backward(self, *tensors):
        self.saved_tensors = tensors


class SS2DFoldedProjFn(torch.autograd.Function):
    """x_proj -> split -> dt_proj -> SS2DFoldedFn as ONE node (m2net.py:179-182 + :193-221).

    As separate nodes the backward pays, per SS2D, a concatenation of (d dts_low, dB, dC) into d x_dbl (the adjoint of
    torch.split), casts of the scan's fp32 dB / dC to the operand dtype on their way there, and an addition of the two
    gradients of xs2 (from the scan and from x_proj).  Inside one node the scan's fp32 dB / dC are written straight into
    their row-slices of d x_dbl (one converting copy each), d dts_low lands in its slice from the GEMM, and x_proj's input
    gradient is accumulated onto the scan's du by the GEMM itself (beta = 1).  The projections stay library GEMMs
    (L is the long, parallel axis); their weight gradients run on nz_proj_wgrad."""

    @staticmethod
    def forward(ctx, xs2, wx, wdt, As, Ds, dt_bias, z, gamma, beta, H, W, eps, out_dtype, R, N):
        from .proj import proj_wgrad  # noqa: F401  (bound late: proj imports nothing from here)
        bsz, two, D, L = xs2.shape
        C = R + 2 * N
        wxc, wdtc = wx.to(xs2.dtype), wdt.to(xs2.dtype)
        x_dbl = torch.matmul(wxc.unsqueeze(0), xs2)                            # (B, 2, 2C, L): array a -> directions 2a, 2a+1
        x4 = x_dbl.view(bsz, 4, C, L)
        dts = torch.matmul(wdtc.unsqueeze(0), x4[:, :, :R])                     # (B, 4, D, L)
        core = _CoreCtx(any(ctx.needs_input_grad[:6]))
        out = SS2DFoldedFn.forward(core, xs2, dts, As, x4[:, :, R:R + N], x4[:, :, R + N:], Ds, dt_bias, z, gamma, beta,
                                   H, W, eps, out_dtype)
        # the scan's dB / dC stay fp32 until they are copied into d x_dbl (SelectiveScanFn.backward casts to in_dtypes)
        t = list(core.inner.in_dtypes)
        t[2] = t[3] = torch.float32
        core.inner.in_dtypes = tuple(t)
        ctx.core = core
        ctx.save_for_backward(xs2, x_dbl, wxc, wdtc)
        ctx.meta = (R, N, wx.dtype, wdt.dtype)
        return out

    @staticmethod
    def backward(ctx, dout):
        from .proj import proj_wgrad
        xs2, x_dbl, wxc, wdtc = ctx.saved_tensors
        R, N, t_wx, t_wdt = ctx.meta
        bsz, two, D, L = xs2.shape
        C = R + 2 * N
        du, ddelta, dA, dB, dC, dD, dbias, dz, dg, db, *_ = SS2DFoldedFn.backward(ctx.core, dout)
        ctx.core = None
        x4 = x_dbl.view(bsz, 4, C, L)
        ddelta = ddelta.view(bsz, 4, D, L)
        dx4 = torch.empty_like(x4)
        dx4[:, :, :R] = torch.matmul(wdtc.transpose(1, 2).unsqueeze(0), ddelta)  # d dts_low            (m2net.py:182)
        dx4[:, :, R:R + N].copy_(dB)                                           # fp32 -> operand dtype, in place   (:181)
        dx4[:, :, R + N:].copy_(dC)
        d_wdt = proj_wgrad(ddelta, x4[:, :, :R]).to(t_wdt) if ctx.needs_input_grad[2] else None
        dx_dbl = dx4.view(bsz, 2, 2 * C, L)
        d_wx = proj_wgrad(dx_dbl, xs2).to(t_wx) if ctx.needs_input_grad[1] else None
        d_xs2 = None
        if ctx.needs_input_grad[0]:                                            # du + W_x^T d x_dbl in one GEMM (:179)
            wT = wxc.transpose(1, 2).unsqueeze(0).expand(bsz, 2, D, 2 * C).reshape(bsz * 2, D, 2 * C)
            d_xs2 = torch.baddbmm(du.reshape(bsz * 2, D, L), wT, dx_dbl.view(bsz * 2, 2 * C, L)).view(bsz, 2, D, L)
        return (d_xs2, d_wx, d_wdt, dA, dD, dbias, dz, dg, db, None, None, None, None, None, None)


def ss2d_core_folded_proj(xs2, wx, wdt, As, Ds, dt_bias, z, gamma, beta, H, W, eps, out_dtype, R, N):
    """xs2 (B, 2, D, L); wx (2, 2 (R + 2N), D): x_proj weights of directions (2a, 2a + 1) stacked per array a; wdt (4, D, R);
    As / Ds / dt_bias in folded direction order."""
    return SS2DFoldedProjFn.apply(xs2, wx, wdt, As, Ds, dt_bias, z, gamma, beta, int(H), int(W), float(eps), out_dtype,
                                  int(R), int(N))


def ss2d_core_folded(xs2, dts, As, Bs, Cs, Ds, dt_bias, z, gamma, beta, H, W, eps, out_dtype):
    """xs2 (B, 2, D, L), dts (B, 4, D, L) contiguous, Bs / Cs (B, 4, N, L) views, all in folded direction order."""
    return SS2DFoldedFn.apply(xs2, dts, As, Bs, Cs, Ds, dt_bias, z, gamma, beta, int(H), int(W), float(eps), out_dtype)


def ss2d_core(xs, dts, As, Bs, Cs, Ds, dt_bias, z, gamma, beta, H, W, eps, out_dtype):
    """xs, dts (B, 4, D, L) contiguous; Bs, Cs (B, 4, N, L) (strided views fine); z (B, H, W, D) -> (B, H, W, D)."""
    return SS2DCoreFn.apply(xs, dts, As, Bs, Cs, Ds, dt_bias, z, gamma, beta, int(H), int(W), float(eps), out_dtype)
