"""``torch.library`` registration of the scan, so that ``torch.compile`` sees one opaque, shape-inferable op.

The reference compiles its networks by default (nnUNetTrainer.py:316-322: ``torch.compile(self.network)`` unless
``nnUNet_compile`` says otherwise).  A ctypes call on ``data_ptr()`` cannot be traced, so under compilation
``selective_scan_fn`` routes through the two custom ops below instead of the plain ``autograd.Function``:

    nnuzoo_b200::scan_fwd(u, delta, A, B, C, D?, z?, delta_bias?, softplus, out_f32, want_fine, rev_mask, u_gdiv)
        -> (out, x, xf)                       xf is an empty tensor when no fine checkpoints were taken
    nnuzoo_b200::scan_bwd(dout, u, delta, A, B, C, D?, z?, delta_bias?, x, xf, softplus, rev_mask, u_gdiv)
        -> (du, ddelta, dA, dB, dC, dD, dz, dbias)      empty tensors stand for the absent optionals

Both have fake (meta) implementations and scan_fwd has an autograd formula, so AOT autograd can trace forward and
backward without running a kernel.  The eager path keeps the ``autograd.Function`` (no dispatcher round trip).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native
from .selective_scan_interface import SelectiveScanFn


class _Ctx:
    """Minimal stand-in for an autograd ctx: lets the ops reuse SelectiveScanFn.forward / .backward verbatim."""

    def __init__(self, want_grad: bool):
        self.needs_input_grad = (want_grad,) * 13
        self.saved_tensors = ()

    def save_for_backward(self, *t):
        self.saved_tensors = t

    def mark_non_differentiable(self, *t):
        pass


def _fine_elems(u, delta, A, B, C, z, want_fine: bool, rev_mask: int, u_gdiv: int) -> int:
    """Number of fp32 fine-checkpoint elements the forward will take for this problem (0: none).  Shape, dtype, stride
    and storage-offset arithmetic only -- the same question SelectiveScanFn.forward asks nz_scan_fine_bytes() with real
    pointers -- so the fake implementation and the kernel launcher agree on the output shape."""
    import ctypes

    from ._native import NzScanDesc
    from .selective_scan_interface import _DTYPES
    if not want_fine or u.dtype not in _DTYPES:
        return 0
    d = NzScanDesc()
    batch, dim, L = delta.shape
    d.batch, d.dim, d.dstate, d.ngroups, d.seqlen = batch, dim, A.shape[1], B.shape[1], L
    d.dtype, d.rev_mask, d.u_gdiv = _DTYPES[u.dtype], int(rev_mask), int(u_gdiv)
    fake = lambda t: ctypes.c_void_p(4096 + t.storage_offset() * t.element_size())  # noqa: E731  (allocations are 512-B aligned)
    d.u, d.delta, d.B, d.C = fake(u), fake(delta), fake(B), fake(C)
    d.u_stride[0], d.u_stride[1] = u.stride(0), u.stride(1)
    d.delta_stride[0], d.delta_stride[1] = delta.stride(0), delta.stride(1)
    if z is not None:
        d.z = fake(z)
        d.z_stride[0], d.z_stride[1] = z.stride(0), z.stride(1)
    for k in range(3):
        d.B_stride[k], d.C_stride[k] = B.stride(k), C.stride(k)
    return int(_native.lib().nz_scan_fine_bytes(ctypes.byref(d))) // 4


def _nchunks(L: int) -> int:
    return (L + _native.NZ_CHUNK - 1) // _native.NZ_CHUNK


@torch.library.custom_op("nnuzoo_b200::scan_fwd", mutates_args=())
def scan_fwd(u: torch.Tensor, delta: torch.Tensor, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor,
             D: Optional[torch.Tensor], z: Optional[torch.Tensor], delta_bias: Optional[torch.Tensor], softplus: bool,
             out_f32: bool, want_fine: bool, rev_mask: int, u_gdiv: int) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    ctx = _Ctx(want_fine)
    out = SelectiveScanFn.forward(ctx, u, delta, A, B, C, D, z, delta_bias, softplus, False,
                                  torch.float32 if out_f32 else None, rev_mask, u_gdiv)
    x = ctx.saved_tensors[8]
    xf = ctx.saved_tensors[9] if ctx.has_xf else torch.empty(0, dtype=torch.float32, device=u.device)
    if xf.numel() != _fine_elems(u, delta, A, B, C, z, want_fine, rev_mask, u_gdiv):
        raise RuntimeError("nnuzoo_b200::scan_fwd: fine-checkpoint size differs from what the fake implementation promised")
    return out, x, xf


@scan_fwd.register_fake
def _(u, delta, A, B, C, D, z, delta_bias, softplus, out_f32, want_fine, rev_mask, u_gdiv):
    batch, dim, L = delta.shape
    out = delta.new_empty((batch, dim, L), dtype=torch.float32 if out_f32 else u.dtype)
    x = delta.new_empty((batch, dim, _nchunks(L), A.shape[1]), dtype=torch.float32)
    xf = delta.new_empty((_fine_elems(u, delta, A, B, C, z, want_fine, rev_mask, u_gdiv),), dtype=torch.float32)
    return out, x, xf


@torch.library.custom_op("nnuzoo_b200::scan_bwd", mutates_args=())
def scan_bwd(dout: torch.Tensor, u: torch.Tensor, delta: torch.Tensor, A: torch.Tensor, B: torch.Tensor, C: torch.Tensor,
             D: Optional[torch.Tensor], z: Optional[torch.Tensor], delta_bias: Optional[torch.Tensor], x: torch.Tensor,
             xf: torch.Tensor, softplus: bool, rev_mask: int, u_gdiv: int
             ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor,
                        torch.Tensor]:
    ctx = _Ctx(True)
    ctx.delta_softplus, ctx.has_z, ctx.has_D, ctx.has_bias = softplus, z is not None, D is not None, delta_bias is not None
    ctx.has_xf = xf.numel() > 0
    ctx.fold = (rev_mask, u_gdiv)
    ctx.squeeze_B = ctx.squeeze_C = False
    ctx.in_dtypes = (delta.dtype, A.dtype, B.dtype, C.dtype, None if D is None else D.dtype,
                     None if z is None else z.dtype, None if delta_bias is None else delta_bias.dtype)
    ctx.saved_tensors = (u, delta, A, B, C, D, z, delta_bias, x) + ((xf,) if ctx.has_xf else ())
    du, ddelta, dA, dB, dC, dD, dz, dbias, *_ = SelectiveScanFn.backward(ctx, dout)
    # (dA / dD / dbias are views of one zero-filled buffer in the eager path; op outputs may not alias each other)
    e = lambda t: t.clone() if t is not None else torch.empty(0, device=u.device)  # noqa: E731
    return du, ddelta, dA.clone(), dB, dC, e(dD), dz if dz is not None else torch.empty(0, device=u.device), e(dbias)


@scan_bwd.register_fake
def _(dout, u, delta, A, B, C, D, z, delta_bias, x, xf, softplus, rev_mask, u_gdiv):
    e = lambda t: torch.empty_like(t) if t is not None else delta.new_empty((0,))  # noqa: E731
    return (torch.empty_like(u), torch.empty_like(delta), torch.empty_like(A), torch.empty_like(B), torch.empty_like(C),
            e(D), e(z), e(delta_bias))


def _setup(ctx, inputs, output):
    u, delta, A, B, C, D, z, delta_bias, softplus, out_f32, want_fine, rev_mask, u_gdiv = inputs
    _, x, xf = output
    ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, x, xf)
    ctx.flags = (softplus, rev_mask, u_gdiv)


def _backward(ctx, dout, _dx, _dxf):
    u, delta, A, B, C, D, z, delta_bias, x, xf = ctx.saved_tensors
    softplus, rev_mask, u_gdiv = ctx.flags
    du, ddelta, dA, dB, dC, dD, dz, dbias = scan_bwd(dout.contiguous(), u, delta, A, B, C, D, z, delta_bias, x, xf, softplus,
                                                     rev_mask, u_gdiv)
    return (du, ddelta, dA, dB, dC, dD if D is not None else None, dz if z is not None else None,
            dbias if delta_bias is not None else None, None, None, None, None, None)


scan_fwd.register_autograd(_backward, setup_context=_setup)


def selective_scan_compiled(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                            return_last_state=False, out_dtype=None):
    """The traceable twin of ``selective_scan_fn`` (same arguments and returns)."""
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    out_f32 = out_dtype == torch.float32 and u.dtype != torch.float32
    want = torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in (u, delta, A, B, C, D, z, delta_bias))
    Df = None if D is None else D.float()
    bf = None if delta_bias is None else delta_bias.float()
    out, x, _ = scan_fwd(u.contiguous(), delta.to(u.dtype).contiguous(), A.float().contiguous(), B.to(u.dtype),
                         C.to(u.dtype), Df, None if z is None else z.to(u.dtype).contiguous(), bf, bool(delta_softplus),
                         bool(out_f32), bool(want), 0, 1)
    if not return_last_state:
        return out
    return out, x[:, :, -1, :].detach()
