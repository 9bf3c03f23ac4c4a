#!/usr/bin/env python
"""CUDA-event timing of the SS2D glue kernels at M2Net (batch 12, 512x512) shapes against the HBM roofline.

    python tools/prof_glue.py            # prints one line per kernel / shape: ms, algorithmic GB/s, % of peak
Algorithmic bytes: proj_wgrad reads G and X once ((M + N) * B*K*L elements); LayerNorm fwd reads x and writes y,
bwd reads dy and x and writes dx (+ 8 bytes of statistics per row); dwconv fwd reads x / writes y, bwd reads x, dy / writes dx.  Each call works on fresh > L2 operands.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from nnuzoo_b200.dwconv import dwconv3x3_silu  # noqa: E402
from nnuzoo_b200.norm import LayerNormFn  # noqa: E402
from nnuzoo_b200.proj import proj_wgrad  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


ONCE = bool(os.environ.get("NZ_PROF_ONCE"))   # one launch per kernel and shape: for `ncu --set full`


def timed(fn, n=10):
    if ONCE:
        n = 1
    for _ in range(0 if ONCE else 3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    pk = peak()
    B, K = 12, 4
    print(f"HBM peak used: {pk:.0f} GB/s")
    # stage: (d_inner, R, L) of the full-resolution SS2D block of each M2Net stage
    for D, R, L in [(32, 1, 512 * 512), (64, 2, 256 * 256), (128, 4, 128 * 128), (256, 8, 64 * 64)]:
        C = R + 32
        for name, M, N in (("x_proj", C, D), ("dt_proj", D, R)):
            g = torch.randn(B, K, M, L, device="cuda").bfloat16()
            x = torch.randn(B, K, N, L, device="cuda").bfloat16()
            ms = timed(lambda: proj_wgrad(g, x))
            nbytes = 2 * (M + N) * B * K * L
            print(f"proj_wgrad {name:7s} D={D:3d} L={L:6d} M={M:3d} N={N:3d}: {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s "
                  f"{100 * nbytes / ms / 1e6 / pk:5.1f} % of peak")
            del g, x
    for C, rows, din, dout in [(16, 12 * 512 * 512, torch.float32, torch.bfloat16),
                               (32, 12 * 512 * 512, torch.float32, torch.float32),
                               (64, 12 * 256 * 256, torch.bfloat16, torch.bfloat16),
                               (128, 12 * 128 * 128, torch.float32, torch.float32),
                               (1024, 12 * 32 * 32, torch.bfloat16, torch.bfloat16)]:
        x = torch.randn(rows, C, device="cuda").to(din).requires_grad_(True)
        w = torch.ones(C, device="cuda", requires_grad=True)
        b = torch.zeros(C, device="cuda", requires_grad=True)
        gy = torch.randn(rows, C, device="cuda").to(dout)
        with torch.no_grad():
            ms_f = timed(lambda: LayerNormFn.apply(x, w, b, 1e-5, dout))
        y = LayerNormFn.apply(x, w, b, 1e-5, dout)
        ms_fb = timed(lambda: torch.autograd.grad(y, (x, w, b), gy, retain_graph=True))
        si, so = x.element_size(), gy.element_size()
        bf = rows * (C * (si + so) + 8)
        bb = rows * (C * (2 * si + so) + 8)
        print(f"layernorm C={C:4d} rows={rows:8d} {str(din)[6:]:>8s}->{str(dout)[6:]:<8s}: fwd {ms_f:6.3f} ms {bf / ms_f / 1e6:6.0f} GB/s "
              f"({100 * bf / ms_f / 1e6 / pk:4.1f} %)  bwd {ms_fb:6.3f} ms {bb / ms_fb / 1e6:6.0f} GB/s ({100 * bb / ms_fb / 1e6 / pk:4.1f} %)")
        del x, gy, y
    for D, res in [(32, 512), (64, 256), (128, 128), (256, 64)]:
        x = torch.randn(B, D, res, res, device="cuda").bfloat16().requires_grad_(True)
        w = torch.randn(D, 1, 3, 3, device="cuda", requires_grad=True)
        b = torch.zeros(D, device="cuda", requires_grad=True)
        gy = torch.randn_like(x)
        with torch.no_grad():
            ms_f = timed(lambda: dwconv3x3_silu(x, w, b))
        y = dwconv3x3_silu(x, w, b)
        ms_b = timed(lambda: torch.autograd.grad(y, (x, w, b), gy, retain_graph=True))
        n = x.numel() * 2
        print(f"dwconv3x3+silu D={D:3d} {res}x{res} bf16: fwd {ms_f:6.3f} ms {2 * n / ms_f / 1e6:6.0f} GB/s "
              f"({100 * 2 * n / ms_f / 1e6 / pk:4.1f} %)  bwd {ms_b:6.3f} ms {3 * n / ms_b / 1e6:6.0f} GB/s "
              f"({100 * 3 * n / ms_b / 1e6 / pk:4.1f} %)")
        del x, gy, y


if __name__ == "__main__":
    main()
