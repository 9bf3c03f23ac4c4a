#!/usr/bin/env python
"""Which collective limits a DDP training step: torch.profiler over ONE M2Net step on every rank (rank 0 reports).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/prof_ddp.py [--sync-bn 1]

For the step it prints: device-busy time, summed NCCL kernel time by kind (gradient all-reduce buckets, SyncBatchNorm
all-gathers, batch-dice all-gather / all-reduce), and the EXPOSED NCCL time = time during which an NCCL kernel runs and
no compute kernel does (interval arithmetic over the kernel timeline).  Not a benchmark: profiler on.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

from nnuzoo_b200.m2net import get_m2net  # noqa: E402
from nnuzoo_b200.train import Trainer, synthetic_batch  # noqa: E402


def union(iv):
    iv = sorted(iv)
    out = []
    for a, b in iv:
        if out and a <= out[-1][1]:
            out[-1][1] = max(out[-1][1], b)
        else:
            out.append([a, b])
    return out


def length(iv):
    return sum(b - a for a, b in iv)


def subtract(a, b):
    """Total length of the union a minus the union b (both sorted, disjoint)."""
    tot, j = 0.0, 0
    for s, e in a:
        cur = s
        while j < len(b) and b[j][1] <= cur:
            j += 1
        k = j
        while k < len(b) and b[k][0] < e:
            if b[k][0] > cur:
                tot += b[k][0] - cur
            cur = max(cur, b[k][1])
            k += 1
        if cur < e:
            tot += e - cur
    return tot


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=12)
    ap.add_argument("--sync-bn", type=int, default=1)
    a = ap.parse_args()
    rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    tr = Trainer(get_m2net(1, 4, True).train(), dev, sync_bn=bool(a.sync_bn))
    data, tg = synthetic_batch(a.batch, 1, 4, patch=(512, 512))
    for _ in range(3):
        tr.train_step(data, tg)
    torch.cuda.synchronize()
    dist.barrier()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        tr.train_step(data, tg)
        torch.cuda.synchronize()
    if rank == 0:
        nccl, comp, kinds = [], [], {}
        for ev in prof.events():
            if ev.device_type != torch.autograd.DeviceType.CUDA:
                continue
            t0 = ev.time_range.start
            t1 = ev.time_range.end
            if "nccl" in ev.name.lower():
                nccl.append((t0, t1))
                kind = ("AllReduce" if "AllReduce" in ev.name else "AllGather" if "AllGather" in ev.name else
                        "ReduceScatter" if "ReduceScatter" in ev.name else "Broadcast" if "Broadcast" in ev.name else "other")
                k = kinds.setdefault(kind, [0.0, 0])
                k[0] += t1 - t0
                k[1] += 1
            elif "memcpy" not in ev.name.lower() and "memset" not in ev.name.lower():
                comp.append((t0, t1))
        un, uc = union(nccl), union(comp)
        span = max(e for _, e in un + uc) - min(s for s, _ in un + uc)
        print(f"world {dist.get_world_size()}, per-GPU batch {a.batch}, sync_bn {bool(a.sync_bn)}: step span {span / 1e3:.1f} ms, "
              f"compute-busy {length(uc) / 1e3:.1f} ms, NCCL-busy {length(un) / 1e3:.1f} ms, "
              f"EXPOSED NCCL (no compute kernel running) {subtract(un, uc) / 1e3:.1f} ms")
        for k, (t, n) in sorted(kinds.items(), key=lambda kv: -kv[1][0]):
            print(f"  ncclDevKernel {k:14s} x{n:<4d} {t / 1e3:8.2f} ms summed")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
