#!/usr/bin/env python
"""Where one M2Net training step spends its GPU time (torch.profiler, kernel level).

    python tools/prof_train.py [--batch 12] [--res 512] > gpurun_out/prof_train.txt

Prints the top kernels by total device time and the share of our own kernels (nz::*), so the next
fusion target is picked from measurements.  Not a benchmark: profiler on.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from nnuzoo_b200.m2net import get_m2net  # noqa: E402
from nnuzoo_b200.train import Trainer, synthetic_batch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--res", type=int, default=512)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    tr = Trainer(get_m2net(1, 4, True).train(), dev)
    data, tg = synthetic_batch(a.batch, 1, 4, patch=(a.res, a.res))
    for _ in range(2):
        tr.train_step(data, tg)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        tr.train_step(data, tg)
        torch.cuda.synchronize()
    rows = {}
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            r = rows.setdefault(ev.name, [0.0, 0])
            r[0] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
            r[1] += 1
    total = sum(v[0] for v in rows.values())
    ours = sum(v[0] for k, v in rows.items() if "nz::" in k or k.startswith("nz_"))
    print(f"batch {a.batch} res {a.res}: device time {total / 1e3:.1f} ms in {sum(v[1] for v in rows.values())} kernels; "
          f"nz:: kernels {ours / 1e3:.1f} ms = {100 * ours / total:.1f} %")
    for k, v in sorted(rows.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"{v[0] / 1e3:9.2f} ms {100 * v[0] / total:5.1f} % x{v[1]:<5d} {k[:150]}")


if __name__ == "__main__":
    main()
