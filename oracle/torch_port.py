"""PyTorch (CPU) port of the reference's ``selective_scan_ref`` -- TEST INFRASTRUCTURE ONLY.

What it is for: (1) the ``cpu_baseline`` / ``--impl reference`` legs of bench.py, where the
reference's own CPU path (a Python loop of torch ops over L, multi-threaded by torch's intra-op
pool) has to be timed on a box where /root/reference does not exist; (2) autograd cross-checks of
the C adjoint in oracle/scan_oracle.c at sizes where autograd through the verbatim reference is
infeasible (the verbatim loop's select-backward is O(L^2)).

Algorithm and operation order follow nnunetv2/nets/seg_mamba/selective_scan_interface.py:86-152
(real A, variable B/C with an optional group axis); the only structural change is that the
discretised tensors are split along L once (``unbind``) before the time loop, so autograd records
one stack/unbind node instead of L selects.  Pinned against the verbatim reference by
oracle/gen_golden.py (see tests/golden/MANIFEST.json: "torch_port_max_abs_diff").
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def selective_scan_port(u, delta, A, B, C, D=None, z=None, delta_bias=None, delta_softplus=False,
                        return_last_state=False):
    in_dtype = u.dtype
    u = u.float()                                             # :102
    delta = delta.float()                                     # :103
    if delta_bias is not None:
        delta = delta + delta_bias.float().unsqueeze(-1)      # :105
    if delta_softplus:
        delta = F.softplus(delta)                             # :107
    bsz, dim, L = u.shape
    n = A.shape[1]
    B = B.float()                                             # :117
    C = C.float()                                             # :118
    if B.dim() == 3:
        B = B.unsqueeze(1)
    if C.dim() == 3:
        C = C.unsqueeze(1)
    # :128 / :131 -- broadcast the group axis over the dim // G rows it serves
    B = B.repeat_interleave(dim // B.shape[1], dim=1)         # (b, d, n, l)
    C = C.repeat_interleave(dim // C.shape[1], dim=1)
    dA = torch.exp(delta.unsqueeze(-1) * A.float().view(1, dim, 1, n))          # :121 (b, d, l, n)
    dBu = (delta.unsqueeze(-1) * B.transpose(2, 3)) * u.unsqueeze(-1)           # :129 (b, d, l, n)
    dA_t = dA.unbind(2)
    dBu_t = dBu.unbind(2)
    C_t = C.unbind(3)
    x = u.new_zeros(bsz, dim, n)                              # :119
    ys = []
    for i in range(L):                                        # :133
        x = dA_t[i] * x + dBu_t[i]                            # :134
        ys.append((x * C_t[i]).sum(-1))                       # :141
    y = torch.stack(ys, dim=2)                                # :147
    out = y if D is None else y + u * D.float().view(1, dim, 1)   # :148
    if z is not None:
        out = out * F.silu(z)                                 # :150 (z keeps its own dtype there)
    out = out.to(in_dtype)                                    # :151
    return (out, x) if return_last_state else out
