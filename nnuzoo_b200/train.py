"""Batch-parallel training step for the SS2D nets -- the caller of the scan path under DDP.

What is mirrored from the reference (and nothing more of its trainer):
  * global batch -> per-rank batch        nnUNetTrainer._set_batch_size_and_oversample, nnUNetTrainer.py:409-429
  * loss: soft Dice (no background, smooth 1e-5, batch dice with the statistics all-gathered across ranks)
    + cross entropy, wrapped for deep supervision with weights 1/2^i, last one ~0, normalised
                                           nnUNetTrainer.py:455-489, loss/dice.py:58-121, loss/compound_losses.py:8-57,
                                           loss/deep_supervision.py:5-30, utilities/ddp_allgather.py:24-50
  * step: zero_grad, autocast forward + loss, backward, clip_grad_norm_(12), AdamW(1e-4, wd 5e-2, eps 1e-5)
                                           nnUNetTrainer.py:1112-1144, nnUNetTrainerM2Net.py:19-22, 55-66
  * targets at the 7 deep-supervision scales [1, 1, 1/2, ... 1/32]   nnUNetTrainerM2Net.py:49-56

One process per GPU (torchrun); gradients are averaged by DistributedDataParallel over NCCL (NVLink / NVSwitch),
bucketed and overlapped with the backward.  The scan itself has no cross-rank exchange.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.distributed as dist
import torch.nn as nn
import torch.nn.functional as F

M2NET_DS_SCALES = (1.0, 1.0, 0.5, 0.25, 0.125, 0.0625, 0.03125)


def split_global_batch(global_batch: int, world_size: int) -> List[int]:
    """Per-rank batch sizes, the first ``global_batch % world_size`` ranks taking one extra sample
    (nnUNetTrainer.py:420-429). 12 -> [12] / [6, 6] / [3, 3, 3, 3] / [2, 2, 2, 2, 1, 1, 1, 1]."""
    if global_batch < world_size:
        raise ValueError("Cannot run DDP if the batch size is smaller than the number of GPUs")
    base, extra = divmod(global_batch, world_size)
    return [base + (1 if r < extra else 0) for r in range(world_size)]


def deep_supervision_weights(n_outputs: int, ddp: bool) -> List[float]:
    """1/2^i, the last output switched off (1e-6 under DDP so every head gets a gradient), normalised to 1
    (nnUNetTrainer.py:473-485)."""
    w = [1.0 / (2 ** i) for i in range(n_outputs)]
    w[-1] = 1e-6 if ddp else 0.0
    s = sum(w)
    return [v / s for v in w]


class _AllGatherWithGrad(torch.autograd.Function):
    """all_gather whose backward sums the incoming gradients over ranks and returns this rank's slice
    (utilities/ddp_allgather.py:24-50)."""

    @staticmethod
    def forward(ctx, t):
        out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
        dist.all_gather(out, t.contiguous())
        return torch.stack(out, 0)

    @staticmethod
    def backward(ctx, g):
        g = g.contiguous()
        dist.all_reduce(g, op=dist.ReduceOp.SUM)
        return g[dist.get_rank()]


class DiceCELoss(nn.Module):
    """DC_and_CE_loss(weight 1 : 1) with MemoryEfficientSoftDiceLoss(do_bg=False, smooth=1e-5) and plain CE."""

    def __init__(self, batch_dice: bool = True, ddp: bool = False, smooth: float = 1e-5):
        super().__init__()
        self.batch_dice, self.ddp, self.smooth = batch_dice, ddp, smooth

    def forward(self, logits: torch.Tensor, target: torch.Tensor) -> torch.Tensor:
        # target (b, 1, *spatial) integer-valued
        labels = target[:, 0].long()
        ce = F.cross_entropy(logits, labels)
        prob = torch.softmax(logits, 1)
        axes = tuple(range(2, prob.dim()))
        with torch.no_grad():
            onehot = torch.zeros(prob.shape, device=prob.device, dtype=torch.bool).scatter_(1, labels.unsqueeze(1), 1)
            onehot = onehot[:, 1:]
            sum_gt = onehot.sum(axes)
        prob = prob[:, 1:]
        intersect = (prob * onehot).sum(axes)
        sum_pred = prob.sum(axes)
        if self.batch_dice:
            # batch dice: the statistics are summed over the batch -- over this rank's samples first, then over ranks
            intersect, sum_pred, sum_gt = intersect.sum(0), sum_pred.sum(0), sum_gt.sum(0)
            if self.ddp:
                # the reference gathers the three per-sample statistics one by one (dice.py:107-111) and sums afterwards;
                # that needs the same batch size on every rank, which its own split of a global batch does not give
                # (12 over 8 ranks = 2,2,2,2,1,1,1,1: nnUNetTrainer.py:420-429).  Summing locally first makes the
                # gathered tensor (3, classes) on every rank -- the same sum, one all-gather (and one all-reduce in the
                # backward) per head instead of three, and no dependence on the per-rank batch.
                packed = torch.stack((intersect, sum_pred, sum_gt.to(sum_pred.dtype)), 0)
                intersect, sum_pred, sum_gt = _AllGatherWithGrad.apply(packed).sum(0).unbind(0)
        dc = (2 * intersect + self.smooth) / torch.clip(sum_gt + sum_pred + self.smooth, 1e-8)
        return ce - dc.mean()


class DeepSupervisionLoss(nn.Module):
    def __init__(self, loss: nn.Module, weights: Sequence[float]):
        super().__init__()
        self.loss, self.weights = loss, tuple(weights)

    def forward(self, outputs, targets):
        return sum(w * self.loss(o, t) for w, o, t in zip(self.weights, outputs, targets) if w != 0.0)


def synthetic_batch(batch: int, in_ch: int, n_classes: int, patch=(512, 512), scales=M2NET_DS_SCALES, seed=0,
                    pin=True):
    """Host-side synthetic batch of the config-2 shape: data N(0,1) (b, in_ch, H, W) fp32 and one random label
    map (b, 1, H*s, W*s) per deep-supervision scale."""
    g = torch.Generator().manual_seed(seed)
    data = torch.randn(batch, in_ch, *patch, generator=g)
    targets = [torch.randint(0, n_classes, (batch, 1, round(patch[0] * s), round(patch[1] * s)), generator=g,
                             dtype=torch.int16) for s in scales]
    if pin and torch.cuda.is_available():
        data, targets = data.pin_memory(), [t.pin_memory() for t in targets]
    return data, targets


class Trainer:
    """One nnUNetTrainerM2Net-style optimisation step on this rank's share of the batch."""

    def __init__(self, network: nn.Module, device, ddp: bool | None = None, sync_bn: bool = False,
                 autocast_dtype=torch.bfloat16, lr=1e-4, weight_decay=5e-2, batch_dice=True, clip=12.0,
                 cuda_graph: bool = False):
        self.device = torch.device(device)
        self.ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1 if ddp is None else ddp
        if sync_bn and self.ddp:
            network = nn.SyncBatchNorm.convert_sync_batchnorm(network)
        network = network.to(self.device)
        self.module = network
        if self.ddp:
            # 1x1 convolution weights (n, c, 1, 1) receive gradients whose size-1 axes carry channels-last strides; DDP
            # then warns "Grad strides do not match bucket view strides" and copies every such gradient into its bucket
            # instead of producing it there.  Normalise the strides where the gradient is born.
            for prm in network.parameters():
                if prm.dim() == 4 and tuple(prm.shape[2:]) == (1, 1):
                    shape, strides = tuple(prm.shape), (prm.shape[1], 1, 1, 1)
                    prm.register_hook(lambda g, sh=shape, st=strides: g if g.stride() == st else (
                        g.as_strided(sh, st) if g.stride()[:2] == st[:2] else g.contiguous().as_strided(sh, st)))
            ids = [self.device.index] if self.device.type == "cuda" else None
            # static_graph: the unused 1x1 heads inside every MU (m2net.py:432) are the same every step
            network = nn.parallel.DistributedDataParallel(network, device_ids=ids, static_graph=True,
                                                          gradient_as_bucket_view=True)
        self.network = network
        n_out = 7 if getattr(self.module, "deep_supervision", False) else 1
        base = DiceCELoss(batch_dice=batch_dice, ddp=self.ddp)
        self.loss = DeepSupervisionLoss(base, deep_supervision_weights(n_out, self.ddp)) if n_out > 1 else base
        # cuda_graph: the whole step (forward, loss, backward, clip, AdamW) is captured once into a CUDA graph and replayed:
        # an M2Net step is ~12 500 kernels and 80 SS2D blocks of ~1.5 ms host time each, so the eager step is bound by the
        # host, not by the GPU.  Static shapes only (nnU-Net patches are); the first `graph_warmup` steps run eagerly.
        self.cuda_graph = bool(cuda_graph) and self.device.type == "cuda" and autocast_dtype != torch.float16
        self._graph, self._static, self._steps, self.graph_warmup, self.graph_error = None, None, 0, 3, None
        self.optimizer = torch.optim.AdamW(self.module.parameters(), lr=lr, weight_decay=weight_decay, eps=1e-5,
                                           betas=(0.9, 0.999), capturable=self.cuda_graph)
        self.autocast_dtype, self.clip = autocast_dtype, clip
        # fp16 autocast needs the reference's loss scaling (nnUNetTrainer.py:1128-1139); bf16 does not
        self.scaler = (torch.amp.GradScaler(self.device.type) if autocast_dtype == torch.float16 and
                       self.device.type == "cuda" else None)

    def _eager_step(self, data, target) -> torch.Tensor:
        dev = self.device
        self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast(dev.type, dtype=self.autocast_dtype, enabled=dev.type == "cuda"):
            out = self.network(data)
            loss = self.loss(out, target)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(self.module.parameters(), self.clip)
        self.optimizer.step()
        return loss.detach()

    def _graphed_step(self, data, target) -> torch.Tensor:
        """Replay (after capturing once) the whole optimisation step on static input buffers."""
        targets = list(target) if isinstance(target, (list, tuple)) else [target]
        if self._graph is None:
            self._static = (torch.empty(data.shape, dtype=data.dtype, device=self.device),
                            [torch.empty(t.shape, dtype=t.dtype, device=self.device) for t in targets])
        s_data, s_targets = self._static
        s_data.copy_(data, non_blocking=True)
        for dst, src in zip(s_targets, targets):
            dst.copy_(src, non_blocking=True)
        if self._graph is None:
            tgt = s_targets if isinstance(target, (list, tuple)) else s_targets[0]
            torch.cuda.synchronize(self.device)
            graph = torch.cuda.CUDAGraph()
            self.optimizer.zero_grad(set_to_none=True)
            with torch.cuda.graph(graph):
                self._static_loss = self._eager_step(s_data, tgt)
            self._graph = graph
        self._graph.replay()
        return self._static_loss

    def train_step(self, data: torch.Tensor, target) -> torch.Tensor:
        """``data`` / ``target`` may be (pinned) host tensors; returns the detached loss on the device."""
        dev = self.device
        self._steps += 1
        if self.cuda_graph and self._steps > self.graph_warmup and self.graph_error is None:
            try:
                return self._graphed_step(data, target)   # copies the (host) batch straight into the static buffers
            except Exception as e:  # capture refused (an op that synchronises, a collective that cannot be captured...)
                if self._graph is not None:
                    raise
                self.graph_error = f"{type(e).__name__}: {e}"[:300]
                torch.cuda.synchronize(dev)
        data = data.to(dev, non_blocking=True)
        if isinstance(target, (list, tuple)):
            target = [t.to(dev, non_blocking=True) for t in target]
        else:
            target = target.to(dev, non_blocking=True)
        self.optimizer.zero_grad(set_to_none=True)
        with torch.autocast(dev.type, dtype=self.autocast_dtype, enabled=dev.type == "cuda"):
            out = self.network(data)
            loss = self.loss(out, target)
        if self.scaler is not None:  # scale -> unscale_ -> clip -> step -> update, as the reference does
            self.scaler.scale(loss).backward()
            self.scaler.unscale_(self.optimizer)
            torch.nn.utils.clip_grad_norm_(self.module.parameters(), self.clip)
            self.scaler.step(self.optimizer)
            self.scaler.update()
        else:
            loss.backward()
            torch.nn.utils.clip_grad_norm_(self.module.parameters(), self.clip)
            self.optimizer.step()
        return loss.detach()
