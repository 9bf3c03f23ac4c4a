"""Eager vs CUDA-graph training step of M2Net: same loss trajectory, time per step.

    python tools/train_graph_check.py [batch] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from nnuzoo_b200.m2net import get_m2net  # noqa: E402
from nnuzoo_b200.train import Trainer, synthetic_batch  # noqa: E402


def run(graph, batch, steps):
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    tr = Trainer(get_m2net(1, 4, True).train(), dev, cuda_graph=graph)
    data, targets = synthetic_batch(batch, 1, 4, seed=17)
    losses = []
    for _ in range(5):
        losses.append(float(tr.train_step(data, targets)))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    acc = torch.zeros(steps, device=dev)
    for i in range(steps):
        acc[i] = tr.train_step(data, targets)
    e1.record()
    torch.cuda.synchronize()
    losses += acc.tolist()
    return e0.elapsed_time(e1) / steps, losses, tr.graph_error


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    for graph in (False, True):
        ms, losses, err = run(graph, batch, steps)
        print(f"graph={graph}: {ms:.1f} ms/step  ({batch / ms * 1e3:.1f} patches/s)  error={err}")
        print("   losses:", " ".join(f"{v:.4f}" for v in losses))
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
