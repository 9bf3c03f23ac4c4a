"""Drop-in for the reference's operator layer, backed by the sm_100a kernels.

Mirrors nnunetv2/nets/seg_mamba/selective_scan_interface.py:14-83 (``SelectiveScanFn`` /
``selective_scan_fn``; the models import the same-named function from
``mamba_ssm.ops.selective_scan_interface``): same signature, same argument meaning, same returns
(``out`` in the dtype and shape of ``u``; ``(out, last_state)`` when ``return_last_state``; the
gradient of ``last_state`` is ignored, :79-82), same tuple order of gradients (:69-74).

Where the reference calls ``selective_scan_cuda.fwd`` / ``.bwd`` (:37, :62) this calls
``nz_scan_fwd`` / ``nz_scan_bwd`` of include/nnuzoo_b200.h.  There is no CPU path: CPU tensors
raise, and a missing native library raises at first use.
"""
from __future__ import annotations

import ctypes

import torch

from . import _native
from ._native import NzScanDesc

_DTYPES = {torch.float32: _native.NZ_F32, torch.bfloat16: _native.NZ_BF16, torch.float16: _native.NZ_F16}


def _ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def _stream(device) -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _check_inputs(u, delta, A, B, C, D, z, delta_bias):
    if not u.is_cuda:
        raise RuntimeError("nnuzoo_b200.selective_scan_fn: tensors must live on a CUDA device "
                           "(there is no CPU fallback; the CPU oracle lives under oracle/ for tests only)")
    if u.dtype not in _DTYPES:
        raise TypeError(f"unsupported dtype {u.dtype}")
    if A.is_complex():
        raise NotImplementedError("complex A is not used by nnUZoo and is not implemented")
    if B.dim() < 3 or C.dim() < 3:
        raise NotImplementedError("only input-dependent (variable) B and C are implemented "
                                  "(the only form nnUZoo uses)")
    if u.dim() != 3 or delta.dim() != 3 or (delta.shape[0], delta.shape[2]) != (u.shape[0], u.shape[2]):
        raise ValueError("u and delta must both be (batch, dim, L)")
    if A.dim() != 2 or A.shape[0] != delta.shape[1]:
        raise ValueError("A must be (dim, dstate)")
    if A.shape[1] > _native.NZ_MAX_DSTATE:
        raise NotImplementedError(f"d_state {A.shape[1]} > {_native.NZ_MAX_DSTATE} is not implemented")


def _fill_common(desc, u, delta, A, B, C, D, z, delta_bias, delta_softplus, force_generic, forward=False, xf=None,
                 rev_mask=0, u_gdiv=1):
    batch, dim, L = delta.shape
    desc.rev_mask, desc.u_gdiv = int(rev_mask), int(u_gdiv)
    desc.batch, desc.dim, desc.dstate, desc.ngroups = batch, dim, A.shape[1], B.shape[1]
    desc.seqlen = L
    desc.dtype = _DTYPES[u.dtype]
    desc.delta_softplus = int(bool(delta_softplus))
    desc.force_generic = int(bool(force_generic))
    desc.u, desc.delta, desc.A, desc.B, desc.C = _ptr(u), _ptr(delta), _ptr(A), _ptr(B), _ptr(C)
    desc.D, desc.z, desc.delta_bias = _ptr(D), _ptr(z), _ptr(delta_bias)
    desc.u_stride[0], desc.u_stride[1] = u.stride(0), u.stride(1)
    desc.delta_stride[0], desc.delta_stride[1] = delta.stride(0), delta.stride(1)
    if z is not None:
        desc.z_stride[0], desc.z_stride[1] = z.stride(0), z.stride(1)
    for k in range(3):
        desc.B_stride[k] = B.stride(k)
        desc.C_stride[k] = C.stride(k)
    desc.A_stride = A.stride(0)
    # scratch for the tile tickets and the chained state hand-off; the call zeroes it on the stream.
    # (the caching allocator orders its reuse after this stream's launches)
    # (the _cp size lets few-rows / long-L forwards run chunk-parallel; it equals the base size for other shapes)
    nbytes = _native.workspace_bytes(batch, dim)
    if forward:
        nbytes = max(nbytes, int(_native.lib().nz_scan_workspace_bytes_cp(ctypes.byref(desc))))
    if xf is not None or rev_mask or u_gdiv > 1:  # row-per-lane kernels: room for their chunk aggregates
        if xf is not None:
            desc.xf = _ptr(xf)
        nbytes = max(nbytes, int(_native.lib().nz_scan_workspace_bytes_bwd(ctypes.byref(desc))))
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=u.device)
    desc.workspace, desc.workspace_bytes = _ptr(ws), nbytes
    return ws


_FORCE_GENERIC = False  # tests flip this to exercise the non-TMA loader on TMA-eligible shapes
_USE_FINE = True        # tests / tools flip this to run the warp-scan backward on shapes the row-per-lane one takes


class SelectiveScanFn(torch.autograd.Function):
    """Same contract as the reference's ``SelectiveScanFn`` (selective_scan_interface.py:14-74)."""

    @staticmethod
    def forward(ctx, u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                return_last_state=False, out_dtype=None, rev_mask=0, u_gdiv=1):
        # rev_mask / u_gdiv (NzScanDesc, ABI v4) are the folded-SS2D extension used by nnuzoo_b200.fused: groups that walk
        # the sequence backwards over un-flipped arrays, and u rows shared by u_gdiv consecutive groups
        _check_inputs(u, delta, A, B, C, D, z, delta_bias)
        if u_gdiv > 1 and u.shape[1] * u_gdiv != delta.shape[1]:
            raise ValueError("with u_gdiv = n, u must be (batch, dim / n, L)")
        if u_gdiv <= 1 and u.shape != delta.shape:
            raise ValueError("u and delta must both be (batch, dim, L)")
        ctx.fold = (int(rev_mask), int(u_gdiv))
        out_f32 = out_dtype == torch.float32 and u.dtype != torch.float32
        if out_dtype is not None and out_dtype != u.dtype and not out_f32:
            raise TypeError("out_dtype must be u's dtype or torch.float32")
        ctx.in_dtypes = tuple(None if t is None else t.dtype for t in (delta, A, B, C, D, z, delta_bias))
        # :19-30 -- only the innermost stride has to be 1; outer strides are passed to the kernel
        if u.stride(-1) != 1:
            u = u.contiguous()
        if delta.stride(-1) != 1 or delta.dtype != u.dtype:
            delta = delta.to(u.dtype).contiguous()
        if D is not None:
            D = D.float().contiguous()
        if delta_bias is not None:
            delta_bias = delta_bias.float().contiguous()
        if B.stride(-1) != 1 or B.dtype != u.dtype:
            B = B.to(u.dtype).contiguous()
        if C.stride(-1) != 1 or C.dtype != u.dtype:
            C = C.to(u.dtype).contiguous()
        if z is not None and (z.stride(-1) != 1 or z.dtype != u.dtype):
            z = z.to(u.dtype).contiguous()
        A = A.float()
        if A.stride(-1) != 1:
            A = A.contiguous()
        # :31-36 -- 3-D B/C get a singleton group axis
        ctx.squeeze_B = B.dim() == 3
        ctx.squeeze_C = C.dim() == 3
        if ctx.squeeze_B:
            B = B.unsqueeze(1)
        if ctx.squeeze_C:
            C = C.unsqueeze(1)
        batch, dim, L = delta.shape
        N = A.shape[1]
        if B.shape != (batch, B.shape[1], N, L) or C.shape != B.shape:
            raise ValueError(f"B/C must be (batch, groups, dstate, L); got {tuple(B.shape)} {tuple(C.shape)}")
        if dim % B.shape[1]:
            raise ValueError("dim must be a multiple of the number of B/C groups")

        lib = _native.lib()
        _native.bind_device(u.device.index)
        nchunks = (L + _native.NZ_CHUNK - 1) // _native.NZ_CHUNK
        out = torch.empty((batch, dim, L), dtype=torch.float32 if out_f32 else u.dtype, device=u.device)
        x = torch.empty((batch, dim, nchunks, N), dtype=torch.float32, device=u.device)
        desc = NzScanDesc()
        ws = _fill_common(desc, u, delta, A, B, C, D, z, delta_bias, delta_softplus, _FORCE_GENERIC, forward=True,
                          rev_mask=rev_mask, u_gdiv=u_gdiv)
        # fine checkpoints (h every NZ_FINE steps) feed the row-per-lane backward; only taken when a gradient will be asked
        # for and the problem qualifies (nz_scan_fine_bytes() > 0)
        xf = None
        if _USE_FINE and any(ctx.needs_input_grad[:8]):
            nfine = int(lib.nz_scan_fine_bytes(ctypes.byref(desc)))
            if nfine:
                xf = torch.empty((nfine // 4,), dtype=torch.float32, device=u.device)
                desc.xf = _ptr(xf)
        desc.out_f32 = int(out_f32)
        desc.out = _ptr(out)
        desc.out_stride[0], desc.out_stride[1] = out.stride(0), out.stride(1)
        desc.x = _ptr(x)
        with torch.cuda.device(u.device):
            _native.check(lib.nz_scan_fwd(ctypes.byref(desc), _stream(u.device)), "nz_scan_fwd")
        del ws
        ctx.delta_softplus = bool(delta_softplus)
        ctx.has_z = z is not None
        ctx.has_D = D is not None
        ctx.has_bias = delta_bias is not None
        ctx.has_xf = xf is not None
        ctx.save_for_backward(u, delta, A, B, C, D, z, delta_bias, x, *([xf] if xf is not None else []))
        if not return_last_state:
            return out
        last_state = x[:, :, -1, :]  # :40 (batch, dim, dstate)
        ctx.mark_non_differentiable(last_state)
        return out, last_state

    @staticmethod
    def backward(ctx, dout, *args):
        u, delta, A, B, C, D, z, delta_bias, x = ctx.saved_tensors[:9]
        xf = ctx.saved_tensors[9] if ctx.has_xf else None
        if dout.stride(-1) != 1 or dout.dtype != u.dtype:  # :57-58
            dout = dout.to(u.dtype).contiguous()
        batch, dim, L = delta.shape
        N, G = A.shape[1], B.shape[1]
        dev = u.device
        lib = _native.lib()
        _native.bind_device(dev.index)
        du = torch.empty((batch, dim, L), dtype=u.dtype, device=dev)
        ddelta = torch.empty_like(du)
        dz = torch.empty_like(du) if ctx.has_z else None
        # the three (dim)-shaped accumulators share one zero fill
        acc = torch.zeros((N + 2) * dim, dtype=torch.float32, device=dev)
        dA = acc[:dim * N].view(dim, N)
        dD = acc[dim * N:dim * (N + 1)] if ctx.has_D else None
        dbias = acc[dim * (N + 1):] if ctx.has_bias else None
        desc = NzScanDesc()
        ws = _fill_common(desc, u, delta, A, B, C, D, z, delta_bias, ctx.delta_softplus, _FORCE_GENERIC, xf=xf,
                          rev_mask=ctx.fold[0], u_gdiv=ctx.fold[1])
        # dB / dC are overwritten when every element has a single owner tile, accumulated into (atomics) otherwise:
        # the library says which
        mk = torch.empty if lib.nz_scan_bwd_overwrites_dbc(ctypes.byref(desc)) else torch.zeros
        dB = mk((batch, G, N, L), dtype=torch.float32, device=dev)
        dC = mk((batch, G, N, L), dtype=torch.float32, device=dev)
        desc.x = _ptr(x)
        desc.dout = _ptr(dout)
        desc.dout_stride[0], desc.dout_stride[1] = dout.stride(0), dout.stride(1)
        desc.du, desc.ddelta, desc.dz = _ptr(du), _ptr(ddelta), _ptr(dz)
        desc.dA, desc.dB, desc.dC, desc.dD, desc.ddelta_bias = _ptr(dA), _ptr(dB), _ptr(dC), _ptr(dD), _ptr(dbias)
        with torch.cuda.device(dev):
            _native.check(lib.nz_scan_bwd(ctypes.byref(desc), _stream(dev)), "nz_scan_bwd")
        del ws
        if ctx.fold[1] > 1:  # groups that share u rows: their input gradients add up
            n = ctx.fold[1]
            du = du.view(batch, G // n, n, dim // G, L).sum(dim=2).view(batch, dim // n, L)
        if ctx.squeeze_B:  # :67-68
            dB = dB.squeeze(1)
        if ctx.squeeze_C:
            dC = dC.squeeze(1)
        # gradients leave in the dtype their inputs arrived in
        t_delta, t_A, t_B, t_C, t_D, t_z, t_bias = ctx.in_dtypes
        cast = lambda g, t: g if g is None or g.dtype == t else g.to(t)  # noqa: E731
        return (du, cast(ddelta, t_delta), cast(dA, t_A), cast(dB, t_B), cast(dC, t_C), cast(dD, t_D),
                cast(dz, t_z), cast(dbias, t_bias), None, None, None, None, None)  # order of :69-74


def selective_scan_fn(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                      return_last_state=False, *, out_dtype=None):
    """if return_last_state is True, returns (out, last_state); last_state has shape
    (batch, dim, dstate) and carries no gradient (selective_scan_interface.py:77-83).

    Extension (keyword-only, not in the reference signature): ``out_dtype=torch.float32`` with 16-bit operands
    returns the fp32 result without the operands ever being widened in HBM -- what SS2D under autocast needs, where
    the reference first copies xs / dts / Bs / Cs to fp32 (m2net.py:185-188) to get an fp32 out_y (:200).  The
    arithmetic is the same fp32 arithmetic on the same operand values; the backward reads ``dout`` in the operand
    dtype."""
    if torch.compiler.is_compiling():
        # under torch.compile (the reference's default, nnUNetTrainer.py:316-322) the scan is one opaque custom op with
        # a fake implementation and an autograd formula (nnuzoo_b200/library_ops.py) instead of a graph break
        from .library_ops import selective_scan_compiled
        return selective_scan_compiled(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state, out_dtype)
    return SelectiveScanFn.apply(u, delta, A, B, C, D, z, delta_bias, delta_softplus, return_last_state, out_dtype)
