"""3-D (6-direction) CrossScan / CrossMerge timing at an SSND 3-D shape (ssnd2net.py:249-255, :285-299).

    python tools/prof_cross3d.py [B D Z H W]
Reports CUDA-event time and GB/s of the algorithmic bytes (scan: read x once + write 6 walks; merge: read 6, write 1).
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nnuzoo_b200 import cross_merge, cross_scan  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    B, D, Z, H, W = (int(a) for a in sys.argv[1:6]) if len(sys.argv) >= 6 else (2, 64, 32, 64, 64)
    dev = torch.device("cuda:0")
    for dt in (torch.float32, torch.bfloat16):
        x = torch.randn(B, D, Z, H, W, device=dev).to(dt)
        es = x.element_size()
        E = x.numel()
        t_scan = timeit(lambda: cross_scan(x))
        print(f"cross_scan 3-D {tuple(x.shape)} {dt}: {t_scan:.3f} ms, {7 * E * es / t_scan / 1e6:.0f} GB/s algorithmic")
    oy = torch.randn(B, 6, D, Z * H * W, device=dev, requires_grad=True)
    t_m = timeit(lambda: cross_merge(oy.detach(), (Z, H, W)))
    print(f"cross_merge 3-D fp32 (reference mode): {t_m:.3f} ms, {7 * E * 4 / t_m / 1e6:.0f} GB/s algorithmic")
    y = cross_merge(oy, (Z, H, W))
    g = torch.randn_like(y)

    def bwd():
        oy.grad = None
        y.backward(g, retain_graph=True)

    t_b = timeit(bwd)
    print(f"cross_merge 3-D backward: {t_b:.3f} ms, {7 * E * 4 / t_b / 1e6:.0f} GB/s algorithmic")


if __name__ == "__main__":
    main()
