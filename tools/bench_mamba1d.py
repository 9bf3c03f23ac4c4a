#!/usr/bin/env python
"""BASELINE configs[2] / [3] (the 1-D Mamba nets) on the scan path: parity-test cases, measured here for the record.

    python tools/bench_mamba1d.py > gpurun_out/mamba1d.txt
cfg 3  Alt1DM2Net / LightUMamba 3d_fullres: u (2, 64, 128^3 = 2 097 152), bf16 I/O, z gate, one B/C group.
cfg 4  MambaND2Net: 56 Mamba layers on 75 / 600 tokens, d_inner 192 / 384 / 768 (launch-bound: microseconds per call).
Algorithmic bytes at the selective_scan_fn boundary with the gate: fwd w(3E + 2S) + 2wE, bwd w(5E + 4S) + 2wE.
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from nnuzoo_b200 import Mamba, causal_conv1d_fn, selective_scan_fn  # noqa: E402


def peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def timed(fn, n=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def scan_case(B, D, L, N=16, dtype=torch.bfloat16, G=1):
    g = torch.Generator(device="cuda").manual_seed(0)
    r = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    u, dl, z = (r(B, D, L).to(dtype).requires_grad_(True) for _ in range(3))
    Bm, Cm = (r(B, G, N, L).to(dtype).requires_grad_(True) for _ in range(2))
    A = (-torch.arange(1, N + 1, device="cuda").float().repeat(D, 1)).requires_grad_(True)
    Dp = torch.ones(D, device="cuda", requires_grad=True)
    bias = torch.full((D,), -2.0, device="cuda", requires_grad=True)
    gout = r(B, D, L).to(dtype)
    with torch.no_grad():
        t_f = timed(lambda: selective_scan_fn(u, dl, A, Bm, Cm, Dp, z, bias, True))

    def fb():
        out = selective_scan_fn(u, dl, A, Bm, Cm, Dp, z, bias, True)
        torch.autograd.grad(out, (u, dl, A, Bm, Cm, Dp, z, bias), gout)

    t_fb = timed(fb)
    w = u.element_size()
    E, S = B * D * L, B * G * N * L
    return t_f, t_fb, w * (5 * E + 2 * S), w * (12 * E + 6 * S)


def main():
    pk = peak()
    print(f"HBM peak used: {pk:.0f} GB/s")
    t_f, t_fb, bf, bfb = scan_case(2, 64, 128 ** 3)
    print(f"cfg 3 scan (2, 64, 2097152) bf16 + z: fwd {t_f:.3f} ms {bf / t_f / 1e6:.0f} GB/s ({100 * bf / t_f / 1e6 / pk:.1f} %), "
          f"fwd+bwd {t_fb:.3f} ms {bfb / t_fb / 1e6:.0f} GB/s ({100 * bfb / t_fb / 1e6 / pk:.1f} %)")
    x = torch.randn(2, 64, 128 ** 3, device="cuda").bfloat16().requires_grad_(True)
    w = torch.randn(64, 4, device="cuda", requires_grad=True)
    b = torch.zeros(64, device="cuda", requires_grad=True)
    with torch.no_grad():
        t = timed(lambda: causal_conv1d_fn(x, w, b, activation="silu"))
    print(f"cfg 3 causal_conv1d + SiLU fwd: {t:.3f} ms {2 * x.numel() * 2 / t / 1e6:.0f} GB/s")
    del x
    blk = Mamba(d_model=32).cuda()
    tok = torch.randn(1, 128 ** 3 // 8, 32, device="cuda", requires_grad=True)      # one eighth of the volume per call
    with torch.autocast("cuda", dtype=torch.bfloat16):
        def step():
            y = blk(tok)
            y.float().square().mean().backward()
        t = timed(step, n=3, warm=1)
    print(f"cfg 3 Mamba(d_model 32) block fwd+bwd on {tok.shape[1]} tokens (bf16 autocast): {t:.2f} ms = {tok.shape[1] / t / 1e3:.1f} M tokens/s")
    t_f, t_fb, bf, bfb = scan_case(1, 128, 512 * 512, dtype=torch.float32, G=4)
    print(f"M2Net batch-1 inference, stage-1 scan (1, 128, 262144) fp32 + z, K = 4: fwd {t_f:.3f} ms {bf / t_f / 1e6:.0f} GB/s "
          f"({100 * bf / t_f / 1e6 / pk:.1f} %), fwd+bwd {t_fb:.3f} ms")
    for D, L in [(192, 600), (384, 600), (768, 75)]:
        t_f, t_fb, _, _ = scan_case(2, D, L)
        print(f"cfg 4 scan (2, {D}, {L}) bf16 + z: fwd {t_f * 1e3:.0f} us, fwd+bwd {t_fb * 1e3:.0f} us per call (launch-bound)")


if __name__ == "__main__":
    main()
