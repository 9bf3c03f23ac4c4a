"""Drop-in ``Mamba`` block for the 1-D nets, built on the sm_100a selective scan.

Constructor signature, parameter names, shapes and initialisation follow the reference's vendored
``Mamba`` (nnunetv2/nets/seg_mamba/mamba_simple.py:37-189) so that ``load_state_dict`` of reference
checkpoints works unchanged -- including the ``*_b`` (backward) and ``*_s`` (slice-interleaved)
parameter sets, which the reference creates unconditionally (:135-181).

``forward`` (:191-357) implements what the reference's fused paths compute:
  bimamba_type "none": mamba_inner_fn      = conv1d + SiLU -> x_proj -> dt_proj -> scan(z gate) -> out_proj
               "v2"  : out + flip(out_b)   (:250-281)
               "v3"  : out + flip(out_b) + un-interleave(out_s)   (:213-249, SegMamba's tri-directional block)
Each direction is ONE autograd node, ``nnuzoo_b200.mamba_inner.MambaInnerFn`` (conv -> x_proj -> dt_proj -> scan with
conv / delta recomputed in the backward, everything L-contiguous, B / C read as strided views: the reference's
MambaInnerFn(NoOutProj), selective_scan_interface.py:159-434).  The ``_b`` direction runs over the UN-flipped xz
(anti-causal convolution + reversed-walk scan), so ``xz.flip([-1])`` / ``out_b.flip([-1])`` (:251, :262) are never
materialised; ``reverse=True`` on the module does the same for a whole block (MambaND's reversed layers,
mamba_nd2net.py:638-656).  Convolutions wider than 4 taps take the op-by-op path around cuDNN.  The incremental
``step`` / inference cache (:359-446) is not used by nnUZoo and is not implemented.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F

from .causal_conv1d import causal_conv1d_fn
from .mamba_inner import mamba_inner_fn, mamba_inner_fn_no_out_proj
from .selective_scan_interface import selective_scan_fn


class Mamba(nn.Module):
    def __init__(self, d_model, d_state=16, d_conv=4, expand=2, dt_rank="auto", dt_min=0.001, dt_max=0.1,
                 dt_init="random", dt_scale=1.0, dt_init_floor=1e-4, conv_bias=True, bias=False,
                 use_fast_path=True, layer_idx=None, device=None, dtype=None, bimamba_type="none", nslices=5,
                 extra_directions=True):
        factory_kwargs = {"device": device, "dtype": dtype}
        super().__init__()
        self.d_model, self.d_state, self.d_conv, self.expand = d_model, d_state, d_conv, expand
        self.d_inner = int(self.expand * self.d_model)
        self.dt_rank = math.ceil(self.d_model / 16) if dt_rank == "auto" else dt_rank
        self.use_fast_path, self.layer_idx = use_fast_path, layer_idx
        self.bimamba_type, self.nslices = bimamba_type, nslices
        if bimamba_type not in ("none", "v2", "v3"):
            raise NotImplementedError(f"bimamba_type {bimamba_type!r}")

        self.in_proj = nn.Linear(self.d_model, self.d_inner * 2, bias=bias, **factory_kwargs)
        self.activation = "silu"
        self.act = nn.SiLU()

        def direction(suffix):  # conv1d / x_proj / dt_proj / A_log / D of one scan direction (:72-181)
            conv = nn.Conv1d(self.d_inner, self.d_inner, bias=conv_bias, kernel_size=d_conv, groups=self.d_inner,
                             padding=d_conv - 1, **factory_kwargs)
            x_proj = nn.Linear(self.d_inner, self.dt_rank + self.d_state * 2, bias=False, **factory_kwargs)
            dt_proj = nn.Linear(self.dt_rank, self.d_inner, bias=True, **factory_kwargs)
            setattr(self, "conv1d" + suffix, conv)
            setattr(self, "x_proj" + suffix, x_proj)
            setattr(self, "dt_proj" + suffix, dt_proj)
            return dt_proj

        dt_proj = direction("")
        # dt projection initialised to preserve variance; bias so that softplus(bias) is in [dt_min, dt_max] (:92-112)
        dt_init_std = self.dt_rank ** -0.5 * dt_scale
        if dt_init == "constant":
            nn.init.constant_(dt_proj.weight, dt_init_std)
        elif dt_init == "random":
            nn.init.uniform_(dt_proj.weight, -dt_init_std, dt_init_std)
        else:
            raise NotImplementedError
        dt = torch.exp(torch.rand(self.d_inner, **factory_kwargs) * (math.log(dt_max) - math.log(dt_min))
                       + math.log(dt_min)).clamp(min=dt_init_floor)
        inv_dt = dt + torch.log(-torch.expm1(-dt))
        with torch.no_grad():
            dt_proj.bias.copy_(inv_dt)
        dt_proj.bias._no_reinit = True

        def s4d_real():  # S4D-real initialisation, kept in fp32 (:114-125)
            A = torch.arange(1, self.d_state + 1, dtype=torch.float32, device=device).repeat(self.d_inner, 1).contiguous()
            p = nn.Parameter(torch.log(A))
            p._no_weight_decay = True
            return p

        def skip():
            p = nn.Parameter(torch.ones(self.d_inner, device=device))
            p._no_weight_decay = True
            return p

        self.A_log, self.D = s4d_real(), skip()
        # extra_directions=False gives the parameter set of upstream ``mamba_ssm.Mamba`` (one direction), which is what
        # lm2net.py:14 / mamba_nd2net.py:26 import; the vendored block creates the other two unconditionally
        if extra_directions or bimamba_type != "none":
            self.A_b_log = s4d_real()          # backward direction (:130-154)
            direction("_b")
            self.D_b = skip()
            self.A_s_log = s4d_real()          # slice-interleaved direction (:158-181)
            direction("_s")
            self.D_s = skip()
        self.out_proj = nn.Linear(self.d_inner, self.d_model, bias=bias, **factory_kwargs)

    # what MambaInnerFnNoOutProj.forward computes (selective_scan_interface.py:159-226)
    def _inner(self, xz, conv1d, x_proj, dt_proj, A_log, D, reverse=False):
        if self.d_conv <= 4 and self.use_fast_path:
            return mamba_inner_fn_no_out_proj(xz, conv1d.weight, conv1d.bias, x_proj.weight, dt_proj.weight,
                                              -torch.exp(A_log.float()), None, None, D.float(),
                                              delta_bias=dt_proj.bias.float(), delta_softplus=True, reverse=reverse)
        if reverse:
            return self._inner(xz.flip([-1]), conv1d, x_proj, dt_proj, A_log, D).flip([-1])
        L = xz.shape[-1]
        x, z = xz.chunk(2, dim=1)
        if self.d_conv <= 4:   # causal depthwise conv + SiLU in one kernel (mamba_simple.py:319-324)
            x = causal_conv1d_fn(x, conv1d.weight.squeeze(1), conv1d.bias, "silu")
        else:
            x = self.act(conv1d(x)[..., :L])
        x_dbl = F.linear(x.transpose(1, 2), x_proj.weight)                    # (b, l, R + 2N)
        dt, B, C = torch.split(x_dbl, [self.dt_rank, self.d_state, self.d_state], dim=-1)
        delta = F.linear(dt, dt_proj.weight).transpose(1, 2)                  # (b, d, l), L-contiguous after copy
        B = B.transpose(1, 2).unsqueeze(1).contiguous()                       # (b, 1, N, l)
        C = C.transpose(1, 2).unsqueeze(1).contiguous()
        A = -torch.exp(A_log.float())
        return selective_scan_fn(x, delta.contiguous(), A, B, C, D.float(), z=z,
                                 delta_bias=dt_proj.bias.float(), delta_softplus=True)

    def forward(self, hidden_states, inference_params=None, *, reverse=False):
        """hidden_states: (B, L, D) -> same shape (mamba_simple.py:191-357).

        ``reverse=True`` (keyword-only extension): ``self(hidden_states.flip(1)).flip(1)`` by addressing -- every
        per-token op commutes with the flip, the convolution runs anti-causally and the scan walks backwards."""
        if inference_params is not None:
            raise NotImplementedError("the incremental-decoding cache is not used by nnUZoo and not implemented")
        batch, seqlen, _ = hidden_states.shape
        if reverse and self.bimamba_type != "none":
            return self.forward(hidden_states.flip(1)).flip(1)
        # (b, 2*d_inner, l) with l innermost straight out of the GEMM (:205-212): no transposed copy
        xz = torch.matmul(self.in_proj.weight, hidden_states.transpose(1, 2))
        if self.in_proj.bias is not None:
            xz = xz + self.in_proj.bias.to(xz.dtype).unsqueeze(-1)
        if self.bimamba_type == "none" and self.d_conv <= 4 and self.use_fast_path:   # (:299-313)
            return mamba_inner_fn(xz, self.conv1d.weight, self.conv1d.bias, self.x_proj.weight, self.dt_proj.weight,
                                  self.out_proj.weight, self.out_proj.bias, -torch.exp(self.A_log.float()), None, None,
                                  self.D.float(), delta_bias=self.dt_proj.bias.float(), delta_softplus=True,
                                  reverse=reverse)
        out = self._inner(xz, self.conv1d, self.x_proj, self.dt_proj, self.A_log, self.D, reverse=reverse)
        if self.bimamba_type in ("v2", "v3"):
            # the backward direction over the un-flipped xz; its result is already un-flipped (:250-281)
            out = out + self._inner(xz, self.conv1d_b, self.x_proj_b, self.dt_proj_b, self.A_b_log, self.D_b,
                                    reverse=True)
        if self.bimamba_type == "v3":
            if seqlen % self.nslices:
                raise ValueError("bimamba v3 needs seqlen to be a multiple of nslices")
            ns = self.nslices
            # slice-interleaved order (:229-247): chunk L into ns slices, interleave them, scan, undo
            xz_s = xz.reshape(batch, xz.shape[1], ns, seqlen // ns).transpose(-1, -2).flatten(-2)
            out_s = self._inner(xz_s, self.conv1d_s, self.x_proj_s, self.dt_proj_s, self.A_s_log, self.D_s)
            out = out + out_s.reshape(batch, self.d_inner, seqlen // ns, ns).permute(0, 1, 3, 2).flatten(-2)
        return self.out_proj(out.transpose(1, 2))
