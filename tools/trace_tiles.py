"""Per-tile timeline of one forward scan (ticket, SM, start, wait, end) -- diagnoses the chained
hand-off between consecutive chunks of a row.   python tools/trace_tiles.py batch kd L"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from nnuzoo_b200 import _native  # noqa: E402
import nnuzoo_b200  # noqa: E402


def main():
    batch, kd, L = (int(a) for a in sys.argv[1:4])
    dev = torch.device("cuda:0")
    lib = _native.lib()
    G, N = 4, 16
    g = torch.Generator(device=dev).manual_seed(0)
    u = torch.randn(batch, kd, L, device=dev, generator=g)
    dl = 0.5 * torch.randn(batch, kd, L, device=dev, generator=g)
    B = torch.randn(batch, G, N, L, device=dev, generator=g)
    C = torch.randn(batch, G, N, L, device=dev, generator=g)
    A = -torch.arange(1, N + 1, device=dev).float().repeat(kd, 1).contiguous()
    D = torch.ones(kd, device=dev)
    bias = torch.full((kd,), -2.0, device=dev)
    rows_per_tile, tl = 16, 256
    nrb = batch * G * ((kd // G + rows_per_tile - 1) // rows_per_tile)
    nch = (L + tl - 1) // tl
    ntiles = nrb * nch
    trace = torch.zeros(ntiles * 8, dtype=torch.int64, device=dev)
    for it in range(3):
        lib.nz_debug_set_trace(ctypes.c_void_p(trace.data_ptr() if it == 2 else 0))
        nnuzoo_b200.selective_scan_fn(u, dl, A, B, C, D, None, bias, True)
        torch.cuda.synchronize()
    lib.nz_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(ntiles, 8)
    ticket = t[:, 0] & 0xffffffff
    fast = t[:, 0] >> 32
    t0 = t[:, 2].min()
    start, loop, wait, end = t[:, 2] - t0, t[:, 3] - t0, t[:, 4], t[:, 5] - t0
    dur = end - start
    print(f"tiles {ntiles} (row blocks {nrb} x chunks {nch}); kernel span {end.max() / 1e3:.1f} us; "
          f"fast-path tiles {fast.mean():.2%}")
    print(f"tile duration us: mean {dur.mean() / 1e3:.2f}  p10 {np.percentile(dur, 10) / 1e3:.2f}  "
          f"p50 {np.percentile(dur, 50) / 1e3:.2f}  p90 {np.percentile(dur, 90) / 1e3:.2f}")
    print(f"chain wait per tile us: mean {wait.mean() / 1e3:.2f}  p50 {np.percentile(wait, 50) / 1e3:.2f}  "
          f"p90 {np.percentile(wait, 90) / 1e3:.2f}   prologue (start->loop) mean {(loop - start).mean() / 1e3:.2f}")
    c = ticket // nrb
    for lo in range(0, nch, max(1, nch // 8)):
        m = c == lo
        print(f"  chunk {lo:4d}: start {start[m].mean() / 1e3:8.1f} us  dur {dur[m].mean() / 1e3:6.2f}  wait {wait[m].mean() / 1e3:6.2f}  "
              f"fast {fast[m].mean():.2f}")
    # lag between a tile's start and its predecessor's start / end
    order = np.argsort(ticket)
    st, en = start[order], end[order]
    lag_s = st[nrb:] - st[:-nrb]
    lag_e = st[nrb:] - en[:-nrb]
    print(f"start(tile) - start(pred): mean {lag_s.mean() / 1e3:.2f} us  p10 {np.percentile(lag_s, 10) / 1e3:.2f}  p90 {np.percentile(lag_s, 90) / 1e3:.2f}")
    print(f"start(tile) - end(pred):   mean {lag_e.mean() / 1e3:.2f} us  p10 {np.percentile(lag_e, 10) / 1e3:.2f}  p90 {np.percentile(lag_e, 90) / 1e3:.2f}")
    # hop latency along a chain: when tile c+1 received the state-0 / last-state carry relative to tile c
    r0, rN = t[:, 6][order] - t0, t[:, 7][order] - t0
    ok = (t[:, 6][order][nrb:] > 0) & (t[:, 6][order][:-nrb] > 0)
    if ok.any():
        h0 = (r0[nrb:] - r0[:-nrb])[ok]
        hN = (rN[nrb:] - rN[:-nrb])[ok]
        print(f"hop (carry-received time, tile c+1 minus tile c): state 0 mean {h0.mean() / 1e3:.2f} us p50 {np.percentile(h0, 50) / 1e3:.2f}  "
              f"last state mean {hN.mean() / 1e3:.2f} us p50 {np.percentile(hN, 50) / 1e3:.2f};  "
              f"in-tile state-0 -> last-state span mean {((rN - r0)[t[:, 6][order] > 0]).mean() / 1e3:.2f} us")
    sm = t[:, 1]
    per_sm = np.bincount(sm.astype(int))
    print(f"tiles per SM: min {per_sm[per_sm > 0].min()} max {per_sm.max()}")


if __name__ == "__main__":
    main()
