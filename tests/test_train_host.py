"""Host-side logic of the batch-parallel training step (nnuzoo_b200/train.py) on CPU: the global-batch split, the
deep-supervision loss against the reference's own loss classes, and -- with two gloo processes -- that the
DDP step (batch-dice statistics all-gathered, gradients averaged) equals the single-process full-batch step.
The network in these tests is a small conv stand-in: the SS2D nets have no CPU path by design.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from nnuzoo_b200.train import (DeepSupervisionLoss, DiceCELoss, Trainer, deep_supervision_weights,
                               split_global_batch, synthetic_batch)


def test_split_global_batch_follows_reference_rule():
    assert split_global_batch(12, 1) == [12]
    assert split_global_batch(12, 2) == [6, 6]
    assert split_global_batch(12, 4) == [3, 3, 3, 3]
    assert split_global_batch(12, 8) == [2, 2, 2, 2, 1, 1, 1, 1]      # uneven at 8 GPUs
    assert split_global_batch(13, 4) == [4, 3, 3, 3]
    with pytest.raises(ValueError):
        split_global_batch(3, 4)


def test_deep_supervision_weights():
    w = deep_supervision_weights(7, ddp=False)
    assert w[-1] == 0.0 and abs(sum(w) - 1) < 1e-12 and abs(w[0] / w[1] - 2) < 1e-12
    w = deep_supervision_weights(7, ddp=True)
    assert 0 < w[-1] < 1e-6 and abs(sum(w) - 1) < 1e-12


@pytest.mark.reference
def test_loss_matches_reference_classes():
    from oracle import ref_loader
    dice = ref_loader.load_file("ref_loss_dice", "nnunetv2/training/loss/dice.py")
    rce = ref_loader.load_file("ref_loss_ce", "nnunetv2/training/loss/robust_ce_loss.py")
    comp = ref_loader.load_file("ref_loss_compound", "nnunetv2/training/loss/compound_losses.py")
    dsw = ref_loader.load_file("ref_loss_ds", "nnunetv2/training/loss/deep_supervision.py")
    comp.RobustCrossEntropyLoss = rce.RobustCrossEntropyLoss
    comp.softmax_helper_dim1 = lambda x: torch.softmax(x, 1)
    ref = comp.DC_and_CE_loss({"batch_dice": True, "smooth": 1e-5, "do_bg": False, "ddp": False}, {}, weight_ce=1,
                              weight_dice=1, ignore_label=None, dice_class=dice.MemoryEfficientSoftDiceLoss)
    weights = deep_supervision_weights(3, ddp=False)
    ref = dsw.DeepSupervisionWrapper(ref, weights)
    ours = DeepSupervisionLoss(DiceCELoss(batch_dice=True, ddp=False), weights)
    torch.manual_seed(0)
    outs = [torch.randn(3, 4, s, s, requires_grad=True) for s in (16, 8, 4)]
    tgts = [torch.randint(0, 4, (3, 1, s, s)).float() for s in (16, 8, 4)]
    a = ref(outs, tgts)
    ga = torch.autograd.grad(a, outs[:2])
    b = ours(outs, tgts)
    gb = torch.autograd.grad(b, outs[:2])
    assert abs(float(a.detach()) - float(b.detach())) < 1e-6
    for x, y in zip(ga[:2], gb[:2]):
        assert torch.allclose(x, y, atol=1e-7, rtol=1e-5)


class _TinyNet(nn.Module):
    deep_supervision = False

    def __init__(self):
        super().__init__()
        self.a = nn.Conv2d(1, 8, 3, padding=1)
        self.b = nn.Conv2d(8, 3, 1)
        self.unused = nn.Conv2d(8, 3, 1)          # like the spare seg heads inside MU (m2net.py:432)

    def forward(self, x):
        return self.b(torch.relu(self.a(x)))


def _ddp_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = _TinyNet()
        tr = Trainer(net, "cpu", lr=1e-2)
        assert tr.ddp
        data, tgt = synthetic_batch(4, 1, 3, patch=(16, 16), scales=(1.0,), seed=5, pin=False)
        sizes = split_global_batch(4, world)
        lo = sum(sizes[:rank])
        for _ in range(2):
            loss = tr.train_step(data[lo:lo + sizes[rank]], tgt[0][lo:lo + sizes[rank]])
        if rank == 0:
            q.put(({k: v.numpy().copy() for k, v in net.state_dict().items()}, float(loss)))
    finally:
        dist.destroy_process_group()


def test_ddp_step_equals_full_batch_step_gloo():
    port = 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    sd2, _ = q.get()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    torch.manual_seed(0)
    net = _TinyNet()
    tr = Trainer(net, "cpu", ddp=False, lr=1e-2)
    data, tgt = synthetic_batch(4, 1, 3, patch=(16, 16), scales=(1.0,), seed=5, pin=False)
    for _ in range(2):
        tr.train_step(data, tgt[0])
    for k, v in net.state_dict().items():
        assert torch.allclose(v, torch.from_numpy(sd2[k]), atol=2e-6, rtol=1e-4), k


def _uneven_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        tr = Trainer(_TinyNet(), "cpu", lr=1e-2)
        data, tgt = synthetic_batch(4, 1, 3, patch=(16, 16), scales=(1.0,), seed=5, pin=False)
        sizes = split_global_batch(4, world)          # [2, 1, 1]
        lo = sum(sizes[:rank])
        for _ in range(2):
            loss = tr.train_step(data[lo:lo + sizes[rank]], tgt[0][lo:lo + sizes[rank]])
        q.put((rank, float(loss), float(sum(p.double().sum() for p in tr.module.parameters()))))
    finally:
        dist.destroy_process_group()


def test_ddp_uneven_batch_split_runs_and_keeps_ranks_in_step_gloo():
    """The reference's own split of a global batch is uneven (12 over 8 ranks = 2,2,2,2,1,1,1,1, nnUNetTrainer.py:420-429).
    The batch-dice all-gather must not depend on the per-rank batch size (it hung an 8-GPU run when it did): world 3,
    global batch 4 -> 2, 1, 1 finishes, and every rank ends with the same parameters."""
    assert split_global_batch(12, 8) == [2, 2, 2, 2, 1, 1, 1, 1]
    port = 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_uneven_worker, args=(r, 3, port, q)) for r in range(3)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0, "a rank hung or failed"
    got = sorted(q.get() for _ in range(3))
    assert all(abs(g[2] - got[0][2]) < 1e-9 for g in got)
    assert all(g[1] == g[1] for g in got)   # finite
