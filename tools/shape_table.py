"""Per-shape fwd / bwd time of the 80 bench scans under the default dispatch, and each shape's share of the step.

    python tools/shape_table.py [iters]
"""
import os
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import torch  # noqa: E402

from bench import BATCH, m2net_scan_list  # noqa: E402
from prof_scan import run  # noqa: E402


def main():
    iters = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    cnt = Counter(m2net_scan_list(512))
    rows, tot = [], 0.0
    for (kd, L), n in sorted(cnt.items(), key=lambda kv: -kv[0][0] * kv[0][1] * kv[1]):
        tf, tb = run(BATCH, kd, L, iters, quiet=True)
        torch.cuda.empty_cache()
        rows.append((kd, L, n, tf, tb))
        tot += n * (tf + tb)
    print(f"{'kd':>5} {'L':>7} {'n':>2} {'fwd ms':>8} {'bwd ms':>8} {'fwd clk':>8} {'bwd clk':>8} {'share':>6}")
    for kd, L, n, tf, tb in rows:
        E = BATCH * kd * L
        c = 1e-3 * 148 * 1.965e9 / E
        print(f"{kd:5d} {L:7d} {n:2d} {tf:8.3f} {tb:8.3f} {tf * c:8.2f} {tb * c:8.2f} {100 * n * (tf + tb) / tot:6.1f}")
    print(f"sum over the step: {tot:.1f} ms")


if __name__ == "__main__":
    main()
