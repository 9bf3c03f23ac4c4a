#!/usr/bin/env python
"""One forward + backward of the full-resolution stage-1 SS2D block of M2Net (batch 12, 512x512, d_model 16, bf16
autocast) -- the launch sequence `ncu --set full -k regex:"nz::|scan|epilogue|dwconv|cross" ...` is pointed at.

    ncu --set full --clock-control none -k regex:"scan_|epilogue|dwconv|cross_scan2d|proj_wgrad|layernorm" -c 16 \
        -f -o /tmp/ss2d python tools/prof_ss2d.py && python tools/ncu_brief.py /tmp/ss2d.ncu-rep > profiles/...
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from nnuzoo_b200 import SS2D  # noqa: E402


def main():
    torch.manual_seed(0)
    blk = SS2D(d_model=16).cuda()
    x = torch.randn(12, 512, 512, 16, device="cuda", requires_grad=True)
    for _ in range(int(os.environ.get("NZ_PROF_WARM", "0")) + 1):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = blk(x)
        y.float().square().mean().backward()
    torch.cuda.synchronize()
    print("ok", float(y.float().abs().mean()))


if __name__ == "__main__":
    main()
