"""Time the depthwise causal conv1d (+SiLU) forward / backward (C ABI) on a cfg-3-like shape.

    python tools/prof_conv.py [batch dim L [dtype]]     default: 2 64 2097152 float32 (LightUMamba stage 1)
Algorithmic bytes: forward 2*E*w (read x, write out), backward 3*E*w (read x, dout; write dx)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nnuzoo_b200 import causal_conv1d_fn  # noqa: E402


def main():
    batch, dim, L = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (2, 64, 2097152)
    dt = getattr(torch, sys.argv[4]) if len(sys.argv) > 4 else torch.float32
    dev = torch.device("cuda:0")
    NB = 4  # rotate operands so nothing is L2-warm
    xs = [torch.randn(batch, dim, L, device=dev, dtype=dt).requires_grad_(True) for _ in range(NB)]
    gos = [torch.randn(batch, dim, L, device=dev, dtype=dt) for _ in range(NB)]
    w = torch.randn(dim, 4, device=dev, requires_grad=True)
    b = torch.randn(dim, device=dev, requires_grad=True)
    E, es = batch * dim * L, xs[0].element_size()
    tf = tb = 0.0
    iters = 6
    for it in range(iters + 2):
        x, go = xs[it % NB], gos[it % NB]
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        y = causal_conv1d_fn(x, w, b, "silu")
        e[1].record()
        y.backward(go)
        e[2].record()
        torch.cuda.synchronize()
        x.grad = None
        if it >= 2:
            tf += e[0].elapsed_time(e[1])
            tb += e[1].elapsed_time(e[2])
    tf, tb = tf / iters, tb / iters
    print(f"causal_conv1d {tuple(xs[0].shape)} {dt}: fwd {tf:.3f} ms ({2 * E * es / tf / 1e6:.0f} GB/s)  "
          f"bwd incl. autograd glue {tb:.3f} ms ({3 * E * es / tb / 1e6:.0f} GB/s)")


if __name__ == "__main__":
    main()
