"""Drop-in ``SS2D`` / ``SSND`` modules built on the sm_100a ops.

Module interface, parameter names, shapes and initialisation follow the reference so that
``load_state_dict`` of reference checkpoints works unchanged (checkpoint row):
  SS2D : nnunetv2/nets/m2net.py:39-225 (same algorithm at SwinUMamba.py:90-278)
  SSND : nnunetv2/nets/ssnd2net.py:73-318 (2-D 4-direction and 3-D 6-direction cross-scan)

Differences are confined to *how* forward_core executes:
  * CrossScan / CrossMerge are single bit-exact CUDA kernels (nnuzoo_b200.cross_scan) instead of
    stack / flip / transpose-contiguous chains;
  * the scan is nnuzoo_b200.selective_scan_fn (no mamba_ssm dependency);
  * B/C are handed to the scan as the strided split views they are (no .contiguous());
  * SS2D's depthwise 3x3 conv + SiLU is one kernel each way (nnuzoo_b200.dwconv);
  * SS2D (2-D) runs scan -> CrossMerge -> out_norm -> SiLU(z) gate as one autograd node with a fused epilogue kernel
    (nnuzoo_b200.fused; ``fuse_epilogue = False`` restores the op-by-op path);
  * the two projection einsums run as batched GEMMs whose weight gradient is our own reduction kernel
    (nnuzoo_b200.proj).
"""
from __future__ import annotations

import os

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import fused
from .cross_scan import cross_merge, cross_scan, cross_scan_pair
from .dwconv import dwconv3x3_silu
from .norm import LayerNorm
from .proj import grouped_proj
from .selective_scan_interface import selective_scan_fn


class _ConvOnly(nn.Sequential):
    """Keeps the reference's state-dict key ``convnd.conv.*`` (monai Convolution(conv_only=True),
    ssnd2net.py:110-120) without depending on monai."""

    def __init__(self, spatial_dims, channels, kernel_size, padding, bias, dilation):
        super().__init__()
        conv_t = {2: nn.Conv2d, 3: nn.Conv3d}[spatial_dims]
        self.add_module("conv", conv_t(channels, channels, kernel_size, padding=padding, groups=channels,
                                       bias=bias, dilation=dilation))


class _CrossScanSSM(nn.Module):
    """Parameters and forward_core shared by SS2D (K = 4) and SSND (K = 4 or 6)."""

    fuse_epilogue = True   # SS2D: run scan -> merge -> out_norm -> gate as one autograd node (nnuzoo_b200.fused)
    fold_directions = True  # ... on the folded direction layout: no flipped copies, the scan walks backwards instead
    fuse_projections = os.environ.get("NZ_FUSE_PROJ", "0") == "1"  # x_proj / split / dt_proj inside the same node
    split_in_proj = os.environ.get("NZ_SPLIT_IN_PROJ", "0") == "1"  # SS2D.forward: in_proj as two GEMMs (no chunk / permute copies)

    def _init_ssm(self, d_model, d_state, expand, dt_rank, dt_min, dt_max, dt_init, dt_scale, dt_init_floor,
                  bias, dropout, k, factory_kwargs):
        self.d_model = d_model
        self.d_state = d_state
        self.expand = expand
        self.d_inner = int(expand * d_model)
        self.dt_rank = math.ceil(d_model / 16) if dt_rank == "auto" else dt_rank
        self.k = k
        self.in_proj = nn.Linear(d_model, self.d_inner * 2, bias=bias, **factory_kwargs)
        self.act = nn.SiLU()
        # K independent input projections, stored stacked: (K, R + 2N, d_inner)   (m2net.py:81-88)
        proj = [nn.Linear(self.d_inner, self.dt_rank + 2 * d_state, bias=False, **factory_kwargs) for _ in range(k)]
        self.x_proj_weight = nn.Parameter(torch.stack([p.weight for p in proj], dim=0))
        # K dt projections: weights (K, d_inner, R), biases (K, d_inner)            (m2net.py:90-102)
        dts = [self.dt_init(self.dt_rank, self.d_inner, dt_scale, dt_init, dt_min, dt_max, dt_init_floor,
                            **factory_kwargs) for _ in range(k)]
        self.dt_projs_weight = nn.Parameter(torch.stack([t.weight for t in dts], dim=0))
        self.dt_projs_bias = nn.Parameter(torch.stack([t.bias for t in dts], dim=0))
        self.A_logs = self.A_log_init(d_state, self.d_inner, copies=k, merge=True)  # (K * d_inner, N)
        self.Ds = self.D_init(self.d_inner, copies=k, merge=True)                   # (K * d_inner)
        self.selective_scan = selective_scan_fn
        self.out_norm = LayerNorm(self.d_inner)
        self.out_proj = nn.Linear(self.d_inner, d_model, bias=bias, **factory_kwargs)
        self.dropout = nn.Dropout(dropout) if dropout > 0.0 else None

    @staticmethod
    def dt_init(dt_rank, d_inner, dt_scale=1.0, dt_init="random", dt_min=0.001, dt_max=0.1, dt_init_floor=1e-4,
                **factory_kwargs):
        """m2net.py:113-139: weight ~ U(+-R^-0.5 * scale) (or constant), bias = softplus^-1(dt),
        dt log-uniform in [dt_min, dt_max]."""
        proj = nn.Linear(dt_rank, d_inner, bias=True, **factory_kwargs)
        std = dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(proj.weight, std)
        elif dt_init == "random":
            nn.init.uniform_(proj.weight, -std, std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(d_inner, **factory_kwargs) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        with torch.no_grad():
            proj.bias.copy_(dt + torch.log(-torch.expm1(-dt)))
        proj.bias._no_reinit = True
        return proj

    @staticmethod
    def A_log_init(d_state, d_inner, copies=1, device=None, merge=True):
        """m2net.py:141-156: S4D-real, A = -(1..N) for every row, kept as log in fp32."""
        a_log = torch.log(torch.arange(1, d_state + 1, dtype=torch.float32, device=device)).repeat(d_inner, 1)
        if copies > 1:
            a_log = a_log.unsqueeze(0).repeat(copies, 1, 1)
            if merge:
                a_log = a_log.flatten(0, 1)
        p = nn.Parameter(a_log.contiguous())
        p._no_weight_decay = True
        return p

    @staticmethod
    def D_init(d_inner, copies=1, device=None, merge=True):
        """m2net.py:158-168: skip parameter, ones, fp32."""
        d = torch.ones(d_inner, device=device)
        if copies > 1:
            d = d.unsqueeze(0).repeat(copies, 1)
            if merge:
                d = d.flatten(0, 1)
        p = nn.Parameter(d.contiguous())
        p._no_weight_decay = True
        return p

    def _scan_operands(self, x: torch.Tensor):
        """x (B, D, *spatial) -> xs, dts (B, K, D, L), As (K*D, N), Bs, Cs (B, K, N, L) views  (m2net.py:172-190)."""
        N, R = self.d_state, self.dt_rank
        xs = cross_scan(x)                                                    # (B, K, D, L)
        x_dbl = grouped_proj(xs, self.x_proj_weight)            # einsum "b k d l, k c d -> b k c l", m2net.py:179
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                                  # :181
        dts = grouped_proj(dts, self.dt_projs_weight)           # einsum "b k r l, k d r -> b k d l", :182
        As = -torch.exp(self.A_logs.float()).view(-1, N)                                    # :190
        return xs, dts, As, Bs, Cs

    def forward_core(self, x: torch.Tensor, merge_mode: str = "reference") -> torch.Tensor:
        """x (B, D, *spatial) -> merged y (B, D, L) fp32 (m2net.py:170-206 + :218; ssnd2net.py:239-298)."""
        bsz = x.shape[0]
        spatial = tuple(x.shape[2:])
        K, N = self.k, self.d_state
        xs, dts, As, Bs, Cs = self._scan_operands(x)
        L = xs.shape[-1]
        if self.selective_scan is selective_scan_fn and xs.dtype != torch.float32 and dts.dtype == xs.dtype:
            # autocast: 16-bit operands go to the kernel as they are, the result comes back in fp32 -- the same
            # numbers as the reference's "widen everything, then scan" (:185-191) without the four fp32 copies
            out_y = selective_scan_fn(
                xs.view(bsz, -1, L), dts.contiguous().view(bsz, -1, L), As, Bs, Cs, self.Ds.float().view(-1), z=None,
                delta_bias=self.dt_projs_bias.float().view(-1), delta_softplus=True, return_last_state=False,
                out_dtype=torch.float32).view(bsz, K, -1, L)
            return cross_merge(out_y, spatial, merge_mode)
        out_y = self.selective_scan(
            xs.float().view(bsz, -1, L), dts.contiguous().float().view(bsz, -1, L),        # :185-186
            As,
            Bs.float(), Cs.float(),                                                         # :187-188 (views)
            self.Ds.float().view(-1), z=None,
            delta_bias=self.dt_projs_bias.float().view(-1),
            delta_softplus=True, return_last_state=False,
        ).view(bsz, K, -1, L)
        return cross_merge(out_y, spatial, merge_mode)

    def _fused_core(self, x: torch.Tensor, z: torch.Tensor):
        """Scan + merge + out_norm + gate as one node (nnuzoo_b200.fused); None when the shape is outside its cover."""
        if (not self.fuse_epilogue or self.k != 4 or x.dim() != 4 or not x.is_cuda
                or self.selective_scan is not selective_scan_fn or not fused.supported(self.d_inner)):
            return None
        H, W = x.shape[2:]
        if z.stride(-1) != 1 or z.stride(1) != z.shape[2] * z.stride(2):
            return None
        if self.fold_directions and self.d_inner % 32 == 0 and (H * W * x.element_size()) % 128 == 0:
            return self._folded_core(x, z)
        xs, dts, As, Bs, Cs = self._scan_operands(x)
        if dts.dtype != xs.dtype:
            return None
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else xs.dtype
        return fused.ss2d_core(xs, dts.contiguous(), As, Bs, Cs, self.Ds.float().view(-1),
                               self.dt_projs_bias.float().view(-1), z, self.out_norm.weight, self.out_norm.bias, H, W,
                               self.out_norm.eps, out_dtype)

    def _folded_core(self, x: torch.Tensor, z: torch.Tensor):
        """The fused core on the folded direction layout (nnuzoo_b200.fused.SS2DFoldedFn): the projections of m2net.py:179-182
        are point-wise in L, so they are taken on the un-flipped arrays xs2 = {row-major, column-major} -- one GEMM per
        array with the weights of its forward and its backward walker stacked -- and the scan walks directions 2 / 3
        backwards instead of reading flipped copies.  Same values as `_scan_operands` + `ss2d_core`, fewer passes."""
        N, R, D = self.d_state, self.dt_rank, self.d_inner
        C = R + 2 * N
        bsz, _, H, W = x.shape

        def fold(p, *tail):
            # reference direction k = 2 * backwards + array  ->  folded k' = 2 * array + backwards (fused.FOLD_PERM):
            # a transpose of the (2, 2) factorisation of the direction axis, one small copy
            return p.view(2, 2, *tail).transpose(0, 1)

        xs2 = cross_scan_pair(x)                                                    # (B, 2, D, L)
        wx = fold(self.x_proj_weight, C, D).reshape(2, 2 * C, D)                    # array a: directions a and a + 2
        wdt = fold(self.dt_projs_weight, D, R).reshape(4, D, R)
        As = -torch.exp(fold(self.A_logs.float(), D, N)).reshape(4 * D, N)          # :190
        Ds = fold(self.Ds.float(), D).reshape(-1)
        bias = fold(self.dt_projs_bias.float(), D).reshape(-1)
        out_dtype = torch.get_autocast_dtype("cuda") if torch.is_autocast_enabled("cuda") else xs2.dtype
        if self.fuse_projections:   # x_proj, split, dt_proj, scan, merge, norm, gate: one node (:179-182, :193-221)
            return fused.ss2d_core_folded_proj(xs2, wx, wdt, As, Ds, bias, z, self.out_norm.weight, self.out_norm.bias,
                                               H, W, self.out_norm.eps, out_dtype, R, N)
        x_dbl = grouped_proj(xs2, wx).view(bsz, 4, C, H * W)                        # folded order, m2net.py:179
        dts, Bs, Cs = torch.split(x_dbl, [R, N, N], dim=2)                          # :181
        dts = grouped_proj(dts, wdt)                                                # :182
        if dts.dtype != xs2.dtype:
            return None
        return fused.ss2d_core_folded(xs2, dts.contiguous(), As, Bs, Cs, Ds, bias, z, self.out_norm.weight,
                                      self.out_norm.bias, H, W, self.out_norm.eps, out_dtype)

    def _finish(self, y, z, bsz, spatial):
        y = y.transpose(1, 2).contiguous().view(bsz, *spatial, -1)     # m2net.py:219
        y = self.out_norm(y)                                            # :220
        y = y * F.silu(z)                                               # :221
        out = self.out_proj(y)                                          # :222
        if self.dropout is not None:
            out = self.dropout(out)
        return out


class SS2D(_CrossScanSSM):
    """nnunetv2/nets/m2net.py:39-225.  forward: (B, H, W, d_model) -> (B, H, W, d_model).

    The same class (identical forward, parameter names and shapes) appears as ``SS2D`` in SwinUMamba.py:90-277
    (constructor spells ``expand`` as ``ssm_ratio``, takes ``act_layer`` and swallows ``**kwargs``), SwinUMambaD.py:154-340
    (``**kwargs``) and LightSS2DMambaUNet.py:77-262; their VSSBlocks construct it as ``SS2D(d_model=..., dropout=...,
    d_state=..., **kwargs)`` (SwinUMamba.py:293).  All of those spellings are accepted here, so one class replaces the four."""

    def __init__(self, d_model=96, d_state=16, d_conv=3, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, dropout=0.0, conv_bias=True, bias=False,
                 device=None, dtype=None, ssm_ratio=None, act_layer=None, **kwargs):
        super().__init__()
        if ssm_ratio is not None:          # SwinUMamba.py:96 names the expansion factor ssm_ratio
            expand = ssm_ratio
        if act_layer is not None and act_layer is not nn.SiLU:
            raise NotImplementedError("SS2D: the fused convolution / gate kernels implement SiLU (the only activation "
                                      "nnUZoo constructs it with: SwinUMamba.py:107)")
        fk = {"device": device, "dtype": dtype}
        self.d_conv = d_conv
        self._init_ssm(d_model, d_state, expand, dt_rank, dt_min, dt_max, dt_init, dt_scale, dt_init_floor, bias,
                       dropout, 4, fk)
        self.conv2d = nn.Conv2d(self.d_inner, self.d_inner, kernel_size=d_conv, padding=(d_conv - 1) // 2,
                                groups=self.d_inner, bias=conv_bias, **fk)

    def forward(self, x: torch.Tensor, **kwargs):
        bsz, H, W, _ = x.shape
        if x.is_cuda and self.split_in_proj:
            # m2net.py:211-213 (in_proj, chunk, permute) as two GEMMs on the halves of the weight: the x half comes out
            # channels-first -- (D, d_model) @ (B, d_model, L) -- which is what the convolution and the scan read, and
            # the z half channels-last, which is what the gate reads.  No permuted copy forward, and backward neither
            # the concatenation of (dx, dz) that chunk's adjoint is nor the permute of dx.
            D = self.d_inner
            wgt, b = self.in_proj.weight, self.in_proj.bias
            flat = x.reshape(bsz, H * W, -1)
            xc = torch.matmul(wgt[:D], flat.transpose(1, 2))              # (B, D, L)
            if b is not None:
                xc = xc + b[:D].to(xc.dtype).unsqueeze(-1)
            z = F.linear(x, wgt[D:], None if b is None else b[D:])        # (B, H, W, D)
            x = xc.view(bsz, D, H, W)
        else:
            x, z = self.in_proj(x).chunk(2, dim=-1)                      # m2net.py:211-212
            x = x.permute(0, 3, 1, 2).contiguous()
        c = self.conv2d
        if x.is_cuda and c.kernel_size == (3, 3) and c.padding == (1, 1) and c.dilation == (1, 1):
            x = dwconv3x3_silu(x, c.weight, c.bias)                      # :214-215 as one kernel
        else:
            x = self.act(c(x))
        g = self._fused_core(x, z)
        if g is not None:                                                # :193-206, :218-221 as one node
            out = self.out_proj(g)                                       # :222
            return out if self.dropout is None else self.dropout(out)
        y = self.forward_core(x)
        return self._finish(y, z, bsz, (H, W))


class SSND(_CrossScanSSM):
    """nnunetv2/nets/ssnd2net.py:73-318.  forward: (B, *spatial, d_model) -> same; 2-D or 3-D.

    ``merge_mode="reference"`` (default) reproduces the reference's 3-D merge bit-exactly, including
    that scan directions 2 and 5 never reach the output (ssnd2net.py:291-298)."""

    def __init__(self, spatial_dims: int, factorization_type: str, d_model: int, d_state=16, d_conv=3, expand=2,
                 dt_rank="auto", dt_min=0.001, dt_max=0.1, dt_init="random", dt_scale=1.0, dt_init_floor=1e-4,
                 dropout=0.0, conv_bias=True, bias=False, device=None, dtype=None, dilation=1,
                 merge_mode: str = "reference"):
        super().__init__()
        if factorization_type != "cross-scan" or spatial_dims not in (2, 3):
            raise Exception("Factorization and spatial_dims are not supported!")  # ssnd2net.py:257
        fk = {"device": device, "dtype": dtype}
        self.spatial_dims = spatial_dims
        self.factorization_type = factorization_type
        self.d_conv = d_conv
        self.merge_mode = merge_mode
        self._init_ssm(d_model, d_state, expand, dt_rank, dt_min, dt_max, dt_init, dt_scale, dt_init_floor, bias,
                       dropout, 2 * spatial_dims, fk)
        self.convnd = _ConvOnly(spatial_dims, self.d_inner, d_conv, (d_conv - 1) // 2, conv_bias, dilation).to(device)

    def forward(self, x: torch.Tensor):
        bsz = x.shape[0]
        spatial = tuple(x.shape[1:-1])
        x, z = self.in_proj(x).chunk(2, dim=-1)                          # ssnd2net.py:305-306
        perm = (0, x.dim() - 1) + tuple(range(1, x.dim() - 1))
        x = self.act(self.convnd(x.permute(*perm).contiguous()))         # :308-309
        y = self.forward_core(x, self.merge_mode)
        return self._finish(y, z, bsz, spatial)
