"""cProfile of the host side of SS2D forward + backward on a tiny input (device time negligible).

    python tools/host_profile.py [fold 0|1] [iters]
"""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nnuzoo_b200.ss2d import SS2D  # noqa: E402


def main():
    fold = bool(int(sys.argv[1])) if len(sys.argv) > 1 else True
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m = SS2D(16).to(dev)
    m.fold_directions = fold
    x = torch.randn(1, 32, 32, 16, device=dev, requires_grad=True)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            m(x).float().sum().backward()

    for _ in range(20):
        step()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(iters):
        step()
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("tottime")
    print(f"fold={fold}: per-call microseconds = tottime / {iters} * 1e6")
    st.print_stats(28)


if __name__ == "__main__":
    main()
