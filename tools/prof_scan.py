"""Launch a few fwd+bwd scans of one shape through the C ABI (for ncu / quick timing).

    python tools/prof_scan.py [batch kd L [iters]]
Prints the CUDA-event time of fwd and bwd and the achieved algorithmic GB/s.
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nnuzoo_b200 import _native  # noqa: E402
from nnuzoo_b200._native import NzScanDesc  # noqa: E402


def run(batch, kd, L, iters=3, quiet=False):
    G, N = 4, 16
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    lib = _native.lib()
    _native.bind_device(0)
    g = torch.Generator(device=dev).manual_seed(0)
    rows, bc = batch * kd * L, batch * G * N * L
    # 3 rotating copies of the operands so nothing is L2-warm between launches
    NB = 3
    u = torch.randn(NB, rows, device=dev, generator=g)
    dl = 0.5 * torch.randn(NB, rows, device=dev, generator=g)
    go = torch.randn(NB, rows, device=dev, generator=g)
    Bm = torch.randn(NB, bc, device=dev, generator=g)
    Cm = torch.randn(NB, bc, device=dev, generator=g)
    out, du, dd = (torch.empty(rows, device=dev) for _ in range(3))
    dB, dC = torch.zeros(bc, device=dev), torch.zeros(bc, device=dev)
    A = -torch.arange(1, N + 1, device=dev).float().repeat(kd, 1).contiguous()
    D = torch.ones(kd, device=dev)
    bias = torch.full((kd,), -2.0, device=dev)
    dA, dD, db = torch.zeros(kd, N, device=dev), torch.zeros(kd, device=dev), torch.zeros(kd, device=dev)
    nch = (L + _native.NZ_CHUNK - 1) // _native.NZ_CHUNK
    x = torch.empty(batch * kd * nch * N, device=dev)
    use_fine = os.environ.get("NZ_PROF_FINE", "1") != "0"
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    st = torch.cuda.current_stream()
    sp = ctypes.c_void_p(st.cuda_stream)

    def desc(i):
        d = NzScanDesc()
        d.batch, d.dim, d.dstate, d.ngroups, d.seqlen = batch, kd, N, G, L
        d.dtype, d.delta_softplus = 0, 1
        d.u, d.delta, d.dout, d.B, d.C = p(u[i]), p(dl[i]), p(go[i]), p(Bm[i]), p(Cm[i])
        d.A, d.D, d.delta_bias = p(A), p(D), p(bias)
        for s in (d.u_stride, d.delta_stride, d.out_stride, d.dout_stride):
            s[0], s[1] = kd * L, L
        for s in (d.B_stride, d.C_stride):
            s[0], s[1], s[2] = G * N * L, N * L, L
        d.A_stride = N
        d.out, d.x, d.du, d.ddelta = p(out), p(x), p(du), p(dd)
        d.dA, d.dB, d.dC, d.dD, d.ddelta_bias = p(dA), p(dB), p(dC), p(dD), p(db)
        return d

    # fine checkpoints + the backward's larger scratch when the row-per-lane backward applies
    d0 = desc(0)
    nfine = int(lib.nz_scan_fine_bytes(ctypes.byref(d0))) if use_fine else 0
    xf = torch.empty(max(nfine // 4, 1), device=dev)
    if nfine:
        d0.xf = p(xf)
    wsb = max(int(lib.nz_scan_workspace_bytes_bwd(ctypes.byref(d0))), int(lib.nz_scan_workspace_bytes_cp(ctypes.byref(d0))))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    if not quiet:
        print(f"fine checkpoints: {nfine} bytes, workspace {wsb} bytes")

    E, S = rows, bc
    tf = tb = 0.0
    for it in range(iters + 1):
        d = desc(it % NB)
        d.workspace, d.workspace_bytes = p(ws), wsb
        if nfine:
            d.xf = p(xf)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        dB.zero_()
        dC.zero_()
        e[0].record(st)
        _native.check(lib.nz_scan_fwd(ctypes.byref(d), sp), "fwd")
        e[1].record(st)
        _native.check(lib.nz_scan_bwd(ctypes.byref(d), sp), "bwd")
        e[2].record(st)
        torch.cuda.synchronize()
        if it > 0:
            tf += e[0].elapsed_time(e[1])
            tb += e[1].elapsed_time(e[2])
    tf, tb = tf / iters, tb / iters
    if quiet:
        return tf, tb
    print(f"shape b={batch} kd={kd} L={L}: fwd {tf:.3f} ms ({4 * (3 * E + 2 * S) / tf / 1e6:.0f} GB/s)  "
          f"bwd {tb:.3f} ms ({4 * (5 * E + 4 * S) / tb / 1e6:.0f} GB/s)  "
          f"clk/elt/SM fwd {tf * 1e-3 * 148 * 1.965e9 / E:.2f} bwd {tb * 1e-3 * 148 * 1.965e9 / E:.2f}")


    return tf, tb


def main():
    batch, kd, L = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (12, 128, 65536)
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    run(batch, kd, L, iters)


if __name__ == "__main__":
    main()
