"""Host-side cost of one SS2D forward + backward (tiny input, so device time is negligible): folded vs un-folded core.

    python tools/host_overhead.py [iters]
"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nnuzoo_b200.ss2d import SS2D  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = SS2D(16).to(dev)
    x = torch.randn(1, 32, 32, 16, device=dev, requires_grad=True)
    for fold in (True, False, True, False):
        m.fold_directions = fold
        for _ in range(20):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                m(x).float().sum().backward()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = m(x)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        for _ in range(iters):
            with torch.autocast("cuda", dtype=torch.bfloat16):
                m(x).float().sum().backward()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"fold={fold}: forward {1e6 * (t1 - t0) / iters:.0f} us, forward+backward {1e6 * (t2 - t1) / iters:.0f} us per call")
    _ = y


if __name__ == "__main__":
    main()
