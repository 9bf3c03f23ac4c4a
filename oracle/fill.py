"""Deterministic parameter fill shared by oracle/gen_golden.py (applied to the REFERENCE network) and by the
parity tests (applied to ours) -- TEST INFRASTRUCTURE ONLY.

M2Net has 41 M parameters, far too many to commit as a fixture, so both sides regenerate the same values from
a seed: keys are visited in sorted order and every tensor is drawn from one seeded CPU generator, shaped by
what the key is (so that BatchNorm variances stay positive, A stays negative, ...).  The fixture then only
stores the input, the outputs and gradient digests.
"""
from __future__ import annotations

import math

import torch


def deterministic_fill(module: torch.nn.Module, seed: int) -> None:
    gen = torch.Generator(device="cpu").manual_seed(seed)
    sd = module.state_dict()
    new = {}
    for key in sorted(sd):
        t = sd[key]
        if not t.is_floating_point():
            new[key] = t.clone()
            continue
        g = torch.randn(t.shape, generator=gen, dtype=torch.float32)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "A_logs":
            n = t.shape[-1]
            v = torch.log(torch.arange(1, n + 1, dtype=torch.float32)).expand_as(g) + 0.1 * g
        elif leaf == "Ds":
            v = 1.0 + 0.1 * g
        elif leaf == "dt_projs_bias":
            v = -4.0 + 0.5 * g
        elif leaf == "running_var":
            v = 1.0 + 0.1 * g.abs()
        elif leaf == "running_mean":
            v = 0.1 * g
        elif t.dim() == 1 and leaf == "weight":       # LayerNorm / BatchNorm scales
            v = 1.0 + 0.1 * g
        elif leaf == "bias":
            v = 0.05 * g
        else:                                          # Linear / Conv / stacked projection weights
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            if leaf in ("x_proj_weight", "dt_projs_weight"):
                fan_in = t.shape[-1]
            v = g / math.sqrt(max(fan_in, 1))
        new[key] = v.to(t.dtype)
    module.load_state_dict(new, strict=True)
