"""The 1-D Mamba block's fused inner function: conv1d + SiLU -> x_proj -> dt_proj -> selective scan (z gate) [-> out_proj]
as ONE autograd node that recomputes the convolution and delta in the backward.

Mirrors ``mamba_inner_fn`` / ``MambaInnerFn`` (nnunetv2/nets/seg_mamba/selective_scan_interface.py:292-434) and
``mamba_inner_fn_no_out_proj`` / ``MambaInnerFnNoOutProj`` (:159-289): same positional arguments, same meaning, same
gradients.  Differences are layout only -- everything stays L-contiguous so that no tensor is ever transposed, flipped or
made contiguous between the kernels:

  * x_dbl is produced TRANSPOSED, (batch, R + 2N, L) = x_proj_weight @ conv_out, so delta's operand and B and C are
    row-slices of it with unit innermost stride: the scan reads B / C through their strides (the reference rearranges and
    ``.contiguous()``-copies both, :336/:345) and delta = delta_proj_weight @ x_dbl[:, :R] lands as (batch, d, L).
  * ``reverse=True`` (keyword-only extension) is the same block applied to the L-flipped sequence with the result flipped
    back -- the ``_b`` direction of the bi-/tri-directional blocks (mamba_simple.py:250-262) and MambaND's reversed layers
    (mamba_nd2net.py:638-656) -- done by addressing: anti-causal convolution (NzConv1dDesc::reverse) and a scan that walks
    from t = L-1 down to 0 (NzScanDesc::rev_mask).  No flipped copy of xz or of the result exists.  Shapes the
    reversed-walk kernels do not take (rows that are not whole 128-byte lines, groups not a multiple of 32 rows) fall back
    to flipping xz / the result around the forward-walking kernels.
  * checkpoint_lvl 1 of the reference (:362-363): conv_out and delta are recomputed in the backward, not saved.

The kernels are nz_causal_conv1d_fwd/_bwd and nz_scan_fwd/_bwd (include/nnuzoo_b200.h); the three small projections are
library GEMMs.  CUDA only -- there is no CPU path.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .causal_conv1d import CausalConv1dFn
from .selective_scan_interface import SelectiveScanFn


class _Ctx:
    """Stand-in for an autograd ctx so the node can drive the conv / scan launchers' forward and backward directly."""

    def __init__(self, n_inputs: int = 13, want_grad: bool = True):
        self.needs_input_grad = (want_grad,) * n_inputs
        self.saved_tensors = ()

    def save_for_backward(self, *t):
        self.saved_tensors = t

    def mark_non_differentiable(self, *t):
        pass


def _rev_by_addressing(x, z, n_state, rank) -> bool:
    """Can the reversed direction run over the un-flipped arrays?  (NzScanDesc::rev_mask needs the row-per-lane kernels.)"""
    from .library_ops import _fine_elems
    batch, d, L = x.shape
    if n_state != 16 or d % 32:
        return False
    # stand-ins with the shapes / strides / offsets the real x_dbl and delta will have (allocations are 512-B aligned)
    xd = torch.empty((batch, rank + 2 * n_state, L), dtype=x.dtype, device="meta")
    delta = torch.empty((batch, d, L), dtype=x.dtype, device="meta")
    B = xd[:, rank:rank + n_state].unsqueeze(1)
    C = xd[:, rank + n_state:].unsqueeze(1)
    A = torch.empty((d, n_state), device="meta")
    return _fine_elems(x, delta, A, B, C, z, True, 1, 1) > 0


class MambaInnerFn(torch.autograd.Function):
    """conv1d + SiLU -> x_proj -> dt_proj -> scan(z) -> optional out_proj, one node (selective_scan_interface.py:292-434)."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias,
                A, D, delta_bias, delta_softplus, reverse, has_out_proj):
        if not xz.is_cuda:
            raise RuntimeError("nnuzoo_b200.mamba_inner_fn: CUDA tensors only (no CPU fallback)")
        if torch.is_autocast_enabled("cuda"):  # :308-313
            dt = torch.get_autocast_dtype("cuda")
            xz = xz.to(dt)
            x_proj_weight, delta_proj_weight = x_proj_weight.to(dt), delta_proj_weight.to(dt)
            if has_out_proj:
                out_proj_weight = out_proj_weight.to(dt)
                out_proj_bias = None if out_proj_bias is None else out_proj_bias.to(dt)
        else:
            x_proj_weight, delta_proj_weight = x_proj_weight.to(xz.dtype), delta_proj_weight.to(xz.dtype)
            if has_out_proj:
                out_proj_weight = out_proj_weight.to(xz.dtype)
                out_proj_bias = None if out_proj_bias is None else out_proj_bias.to(xz.dtype)
        if xz.stride(-1) != 1:
            xz = xz.contiguous()
        w = conv1d_weight.reshape(conv1d_weight.shape[0], conv1d_weight.shape[-1]).float().contiguous()  # "d 1 w -> d w"
        cb = None if conv1d_bias is None else conv1d_bias.float().contiguous()
        R, N = delta_proj_weight.shape[1], A.shape[-1]
        x, z = xz.chunk(2, dim=1)
        flip = bool(reverse) and not _rev_by_addressing(x, z, N, R)
        if flip:  # layout fallback: the forward-walking kernels around flipped copies
            xz = xz.flip(-1)
            x, z = xz.chunk(2, dim=1)
        rev = bool(reverse) and not flip
        with torch.autocast("cuda", enabled=False):
            conv_out = CausalConv1dFn.forward(_Ctx(5), x, w, cb, True, rev)
            x_dbl = torch.matmul(x_proj_weight, conv_out)                        # (b, R + 2N, L), L-contiguous
            delta = torch.matmul(delta_proj_weight, x_dbl[:, :R])                # (b, d, L)
            B = x_dbl[:, R:R + N].unsqueeze(1)                                   # (b, 1, N, L) views, unit inner stride
            C = x_dbl[:, R + N:].unsqueeze(1)
            sctx = _Ctx(13, any(ctx.needs_input_grad) or rev)  # fine checkpoints only when a backward will follow
            out = SelectiveScanFn.forward(sctx, conv_out, delta, A, B, C, D, z, delta_bias, delta_softplus, False, None,
                                          int(rev), 1)
            u_s, delta_s, A_s, B_s, C_s, D_s, z_s, bias_s, xck = sctx.saved_tensors[:9]
            xf = sctx.saved_tensors[9] if sctx.has_xf else None
            del conv_out, delta, u_s, delta_s  # recomputed in the backward (checkpoint_lvl 1, :362-363)
            ctx.scan_meta = (sctx.squeeze_B, sctx.squeeze_C, sctx.in_dtypes, sctx.has_xf)
            ctx.flags = (bool(delta_softplus), rev, flip, bool(has_out_proj), out_proj_bias is not None,
                         conv1d_weight.shape, R, N)
            ctx.in_dtypes = tuple(None if t is None else t.dtype for t in
                                  (conv1d_weight, conv1d_bias, A, D, delta_bias))
            ctx.save_for_backward(xz, w, cb, x_dbl, x_proj_weight, delta_proj_weight,
                                  out_proj_weight if has_out_proj else None, A_s, D_s, bias_s, xck, xf,
                                  out if has_out_proj else None)
            if has_out_proj:
                y = F.linear(out.transpose(1, 2), out_proj_weight, out_proj_bias)  # (b, L, d_model)  (:367)
                return y.flip(1) if flip else y
            return out.flip(-1) if flip else out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, dy):
        (xz, w, conv_b, x_dbl, x_proj_w, dt_proj_w, out_proj_w, A, D, bias, xck, xf, out) = ctx.saved_tensors
        softplus, rev, flip, has_out_proj, has_out_bias, conv_w_shape, R, N = ctx.flags
        x, z = xz.chunk(2, dim=1)
        with torch.autocast("cuda", enabled=False):
            d_out_w = d_out_b = None
            if has_out_proj:
                dy = dy.to(out.dtype)
                if flip:
                    dy = dy.flip(1)
                # dout_y (b, d, L) = W_out^T dy^T, L-contiguous (:390-391)
                dout = torch.matmul(out_proj_w.t(), dy.transpose(1, 2))
                d_out_w = torch.einsum("ble,bdl->ed", dy, out)                    # (:397)
                d_out_b = dy.sum(dim=(0, 1)) if has_out_bias else None
            else:
                dout = dy.flip(-1) if flip else dy
            # recompute conv_out and delta (:380-383)
            conv_out = CausalConv1dFn.forward(_Ctx(5), x, w, conv_b, True, rev)
            delta = torch.matmul(dt_proj_w, x_dbl[:, :R])
            B = x_dbl[:, R:R + N].unsqueeze(1)
            C = x_dbl[:, R + N:].unsqueeze(1)
            sctx = _Ctx(13)
            sctx.delta_softplus, sctx.has_z, sctx.has_D, sctx.has_bias = softplus, True, D is not None, bias is not None
            sctx.squeeze_B, sctx.squeeze_C, sctx.in_dtypes, sctx.has_xf = ctx.scan_meta
            sctx.fold = (int(rev), 1)
            sctx.saved_tensors = (conv_out, delta, A, B, C, D, z, bias, xck) + ((xf,) if xf is not None else ())
            dconv, ddelta, dA, dB, dC, dD, dz, dbias, *_ = SelectiveScanFn.backward(sctx, dout)
            # x_dbl's gradient, assembled in its own (b, R + 2N, L) layout (:398-421)
            dx_dbl = torch.empty_like(x_dbl)
            dx_dbl[:, :R] = torch.matmul(dt_proj_w.t(), ddelta)
            dx_dbl[:, R:R + N].copy_(dB[:, 0])
            dx_dbl[:, R + N:].copy_(dC[:, 0])
            d_dt_w = torch.einsum("bdl,brl->dr", ddelta, x_dbl[:, :R])            # (:418)
            d_x_w = torch.einsum("brl,bdl->rd", dx_dbl, conv_out)                 # (:421)
            dconv = torch.baddbmm(dconv, x_proj_w.t().expand(x.shape[0], -1, -1), dx_dbl)  # (:422)
            # convolution backward (:426-428)
            cctx = _Ctx(5)
            cctx.saved_tensors = (x, w, conv_b)
            cctx.silu, cctx.reverse = True, rev
            cctx.in_dtypes = (torch.float32, None if conv_b is None else torch.float32)
            dx, dw, db, *_ = CausalConv1dFn.backward(cctx, dconv)
            dxz = torch.cat([dx, dz], dim=1)
            if flip:
                dxz = dxz.flip(-1)
        t_cw, t_cb, t_A, t_D, t_bias = ctx.in_dtypes
        cast = lambda g, t: g if g is None or t is None or g.dtype == t else g.to(t)  # noqa: E731
        return (dxz, cast(dw.reshape(conv_w_shape), t_cw), cast(db, t_cb), d_x_w, d_dt_w, d_out_w, d_out_b,
                cast(dA, t_A), cast(dD, t_D), cast(dbias, t_bias), None, None, None)


def _no_constant_bc(B, C, B_proj_bias, C_proj_bias):
    if B is not None or C is not None or B_proj_bias is not None or C_proj_bias is not None:
        raise NotImplementedError("only input-dependent B and C without projection biases are implemented "
                                  "(the only form nnUZoo uses: mamba_simple.py:218-262, :299-313)")


def mamba_inner_fn(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight, out_proj_bias, A,
                   B=None, C=None, D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True, *,
                   reverse=False):
    """xz: (batch, 2 * d_inner, L) -> (batch, L, d_model)   (selective_scan_interface.py:609-618)."""
    _no_constant_bc(B, C, B_proj_bias, C_proj_bias)
    return MambaInnerFn.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, out_proj_weight,
                              out_proj_bias, A, D, delta_bias, delta_softplus, reverse, True)


def mamba_inner_fn_no_out_proj(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, A, B=None, C=None,
                               D=None, delta_bias=None, B_proj_bias=None, C_proj_bias=None, delta_softplus=True, *,
                               reverse=False):
    """xz: (batch, 2 * d_inner, L) -> (batch, d_inner, L)   (selective_scan_interface.py:159-289, :621-628)."""
    _no_constant_bc(B, C, B_proj_bias, C_proj_bias)
    return MambaInnerFn.apply(xz, conv1d_weight, conv1d_bias, x_proj_weight, delta_proj_weight, None, None, A, D,
                              delta_bias, delta_softplus, reverse, False)
