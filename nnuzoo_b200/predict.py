"""Sliding-window inference with the tiles sharded over GPUs and a gaussian-weighted logit merge
(BASELINE config 5).

Mirror of the reference predictor's inner path:
  compute_gaussian / compute_steps_for_sliding_window   nnunetv2/inference/sliding_window_prediction.py:11-56
  slicers (2-D tiles walked slice by slice through a volume, or same-dimensional tiles)
                                                         nnunetv2/inference/predict_from_raw_data.py:515-547
  mirror-and-predict test-time augmentation              :549-565
  fp16 accumulators, ``logits += prediction * gaussian``, ``n += gaussian``, final division, inf check
                                                         :567-634
What is different, by design:
  * on a CUDA device the whole per-tile arithmetic -- mirror average (:549-565), conversion to the accumulator dtype,
    gaussian weighting and both accumulations (:617-623) -- is ONE kernel per batch of tiles (``nz_sw_accumulate``,
    csrc/sw_kernels.cu) instead of ~3 + 2 * mirrors element-wise launches per tile; it rounds after every step exactly
    as the PyTorch expressions do, so the accumulators are bit-identical (tests/test_predict_gpu.py).  Overlapping tiles
    of one batch are issued one by one.  ``fused_accumulate=False`` keeps the eager expressions;
  * ``tile_batch`` tiles go through the network per forward (the reference uses batch 1, :614), and the mirrored
    copies of a batch are stacked into the same forward when ``stack_mirrors`` -- an eval-mode network treats samples
    independently, so the per-tile results are the same numbers;
  * the tile list is sharded over the ranks of the default process group.  Disjoint tiles (2-D slices of a volume,
    one tile per slice -- config 5): contiguous balanced runs of slices per rank, merged with ONE all-gather of the
    finished slabs (exact).  Overlapping tiles: ``slicers[rank::world]``, every rank accumulates its own tiles and
    the two accumulators are summed with one all-reduce each (NCCL over NVLink) *before* the division (fp16
    summation-order tolerance).
"""
from __future__ import annotations

import itertools
import math
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


_SW_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _gauss_1d(n: int, sigma: float, truncate: float = 4.0) -> torch.Tensor:
    """Response of scipy.ndimage.gaussian_filter (mode constant) to a unit impulse at n // 2, one axis, fp64."""
    radius = int(truncate * sigma + 0.5)
    x = torch.arange(-radius, radius + 1, dtype=torch.float64)
    w = torch.exp(-0.5 * (x / sigma) ** 2)
    w = w / w.sum()
    out = torch.zeros(n, dtype=torch.float64)
    c = n // 2
    lo, hi = max(0, c - radius), min(n, c + radius + 1)
    out[lo:hi] = w[lo - (c - radius):hi - (c - radius)]
    return out


def compute_gaussian(tile_size: Sequence[int], sigma_scale: float = 1.0 / 8, value_scaling_factor: float = 1.0,
                     dtype=torch.float16, device="cpu") -> torch.Tensor:
    """Importance map of one tile (sliding_window_prediction.py:11-29): a gaussian centred on the tile, sigma =
    size * sigma_scale per axis, scaled to max = value_scaling_factor, zeros replaced by the smallest non-zero."""
    g = None
    for n in tile_size:
        a = _gauss_1d(int(n), n * sigma_scale)
        g = a if g is None else g.unsqueeze(-1) * a
    g = (g / g.max() * value_scaling_factor).to(dtype)
    nz = g != 0
    if not bool(nz.all()):
        g[~nz] = g[nz].min()
    return g.to(device)


def compute_steps_for_sliding_window(image_size: Sequence[int], tile_size: Sequence[int],
                                     tile_step_size: float) -> List[List[int]]:
    """Tile origins per axis: at most tile * step apart, evenly spread, last tile flush with the border
    (sliding_window_prediction.py:32-56)."""
    if not 0 < tile_step_size <= 1:
        raise ValueError("step_size must be larger than 0 and smaller or equal to 1")
    steps = []
    for size, tile in zip(image_size, tile_size):
        if size < tile:
            raise ValueError("image size must be as large or larger than patch_size")
        n = int(math.ceil((size - tile) / (tile * tile_step_size))) + 1
        span = size - tile
        stride = span / (n - 1) if n > 1 else 0.0
        steps.append([int(round(stride * i)) for i in range(n)])
    return steps


def sliding_window_slicers(image_size: Sequence[int], patch_size: Sequence[int], tile_step_size: float):
    """Tuples indexing a (c, *image_size) array, in the reference's order (predict_from_raw_data.py:515-547):
    with a patch one dimension short of the image, every index of the first axis is tiled in 2-D."""
    image_size, patch_size = tuple(image_size), tuple(patch_size)
    slicers = []
    if len(patch_size) < len(image_size):
        if len(patch_size) != len(image_size) - 1:
            raise ValueError("patch_size may be at most one dimension shorter than the image")
        steps = compute_steps_for_sliding_window(image_size[1:], patch_size, tile_step_size)
        for d in range(image_size[0]):
            for origin in itertools.product(*steps):
                slicers.append((slice(None), d, *[slice(o, o + t) for o, t in zip(origin, patch_size)]))
    else:
        steps = compute_steps_for_sliding_window(image_size, patch_size, tile_step_size)
        for origin in itertools.product(*steps):
            slicers.append((slice(None), *[slice(o, o + t) for o, t in zip(origin, patch_size)]))
    return slicers


def pad_to_patch(image: torch.Tensor, patch_size: Sequence[int]):
    """Centre zero-padding of the trailing axes up to the patch size (acvl_utils pad_nd_image as called at
    predict_from_raw_data.py:668-670); returns the padded image and the slicer that undoes it."""
    nd = len(patch_size)
    shape = image.shape[-nd:]
    pads, revert = [], [slice(None)] * (image.dim() - nd)
    for s, p in zip(shape, patch_size):
        total = max(p - s, 0)
        lo = total // 2
        pads.append((lo, total - lo))
        revert.append(slice(lo, lo + s))
    if any(a or b for a, b in pads):
        flat = [v for ab in reversed(pads) for v in ab]
        image = torch.nn.functional.pad(image, flat, mode="constant", value=0)
    return image, tuple(revert)


class SlidingWindowPredictor:
    """``predict_logits((c, *spatial)) -> (heads, *spatial)`` fp16, the reference's
    predict_sliding_window_return_logits (predict_from_raw_data.py:646-690) for one set of weights."""

    def __init__(self, network: torch.nn.Module, patch_size: Sequence[int], num_heads: int, device,
                 tile_step_size: float = 0.5, use_gaussian: bool = True, use_mirroring: bool = True,
                 mirror_axes: Sequence[int] | None = None, tile_batch: int = 4, stack_mirrors: bool = True,
                 autocast_dtype=torch.float16, results_dtype=torch.float16, fused_accumulate: bool = True):
        self.device = torch.device(device)
        self.network = network.to(self.device).eval()
        self.patch_size = tuple(int(p) for p in patch_size)
        self.num_heads = num_heads
        self.tile_step_size = tile_step_size
        self.use_gaussian = use_gaussian
        axes = tuple(range(len(self.patch_size))) if mirror_axes is None else tuple(mirror_axes)
        self.mirror_axes = axes if use_mirroring else ()
        self.tile_batch = max(1, int(tile_batch))
        self.stack_mirrors = stack_mirrors
        self.autocast_dtype = autocast_dtype
        self.results_dtype = results_dtype
        self.fused_accumulate = fused_accumulate
        self.forwards = 0   # network invocations of the last call (for the bench)

    # -- mirror average + gaussian multiply-accumulate of a batch of tiles as one kernel ----------------------------
    def _combos(self):
        return [c for i in range(len(self.mirror_axes))
                for c in itertools.combinations([m + 2 for m in self.mirror_axes], i + 1)]

    def _can_fuse(self, data) -> bool:
        return (self.fused_accumulate and self.device.type == "cuda" and self.stack_mirrors
                and self.results_dtype in _SW_DT and data.dim() - 1 <= 3 and len(self._combos()) + 1 <= 8)

    def _predict_stacked(self, x: torch.Tensor) -> torch.Tensor:
        """All mirrored copies of a batch in one forward; returns the raw stacked output (passes, then tiles)."""
        combos = self._combos()
        self.forwards += 1
        if not combos:
            return self.network(x)
        return self.network(torch.cat([x] + [torch.flip(x, c) for c in combos], 0))

    def _accumulate_fused(self, out, group, logits, n_pred, gaussian, nd_vol):
        import ctypes

        from . import _native
        combos = self._combos()
        n = len(group)
        nt = len(self.patch_size)                      # tile axes are the LAST nt axes of the volume
        lead = nd_vol - nt                             # 2-D tiles walked through a 3-D volume: one leading slice axis
        pad = 3 - nd_vol
        tile = [1] * (pad + lead) + list(self.patch_size)
        vol = [1] * pad + list(logits.shape[1:])
        offs = []
        for sl in group:                               # sl = (slice(None), *per-axis slices or ints)
            o = []
            for s_ in sl[1:]:
                o.append(int(s_.start) if isinstance(s_, slice) else int(s_))
            offs.append([0] * pad + o)
        masks = [0] + [sum(1 << (pad + lead + (ax - 2)) for ax in c) for c in combos]
        if out.dtype not in _SW_DT:
            out = out.float()
        out = out.contiguous()
        lib = _native.lib()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _native.bind_device(dev_index)
        st = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        i64 = lambda v: (ctypes.c_int64 * len(v))(*v)  # noqa: E731
        mk = (ctypes.c_int32 * len(masks))(*masks)

        def launch(first, count):
            # pass m of tile t sits at m * n + t: a sub-range of tiles needs its own compact stack
            src = out if count == n else torch.cat([out[m * n + first:m * n + first + count] for m in range(len(masks))], 0)
            flat = [v for o in offs[first:first + count] for v in o]
            with torch.cuda.device(self.device):
                _native.check(lib.nz_sw_accumulate(
                    ctypes.c_void_p(src.data_ptr()), _SW_DT[src.dtype], len(masks), mk, count, self.num_heads, i64(tile),
                    ctypes.c_void_p(gaussian.data_ptr()), _SW_DT[self.results_dtype], ctypes.c_void_p(logits.data_ptr()),
                    ctypes.c_void_p(n_pred.data_ptr()), i64(vol), i64(flat), st), "nz_sw_accumulate")

        def overlap(a, b):
            return all(a[k] < b[k] + tile[k] and b[k] < a[k] + tile[k] for k in range(3))

        if any(overlap(offs[i], offs[j]) for i in range(n) for j in range(i)):
            for t in range(n):
                launch(t, 1)
        else:
            for first in range(0, n, 32):
                launch(first, min(32, n - first))

    # -- test-time augmentation (predict_from_raw_data.py:549-565) on a batch of tiles ------------------------------
    def _mirror_and_predict(self, x: torch.Tensor) -> torch.Tensor:
        combos = [c for i in range(len(self.mirror_axes))
                  for c in itertools.combinations([m + 2 for m in self.mirror_axes], i + 1)]
        if not combos:
            self.forwards += 1
            return self.network(x)
        if self.stack_mirrors:
            n = x.shape[0]
            out = self.network(torch.cat([x] + [torch.flip(x, c) for c in combos], 0))
            self.forwards += 1
            pred = out[:n].clone()
            for i, c in enumerate(combos, start=1):
                pred += torch.flip(out[i * n:(i + 1) * n], c)
        else:
            pred = self.network(x)
            for c in combos:
                pred += torch.flip(self.network(torch.flip(x, c)), c)
            self.forwards += 1 + len(combos)
        pred /= (len(combos) + 1)
        return pred

    @torch.no_grad()
    def predict_logits(self, image: torch.Tensor) -> torch.Tensor:
        dev = self.device
        ddp = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1
        rank, world = (dist.get_rank(), dist.get_world_size()) if ddp else (0, 1)
        self.forwards = 0
        data, revert = pad_to_patch(image, self.patch_size)
        data = data.to(dev)
        slicers = sliding_window_slicers(data.shape[1:], self.patch_size, self.tile_step_size)
        # Sharding.  2-D tiles of a volume whose slices are covered by ONE tile each are disjoint in the output: every
        # rank then takes a contiguous, balanced run of slices (200 slices on 8 ranks = 25 each; `rank::world` with
        # tile batches of 4 gave 7, 7, 6, ... forwards) and the merge is an all-gather of the finished slabs -- 1/world
        # of the bytes of the all-reduce below and no summation at all.  Everything else (overlapping tiles) keeps
        # `rank::world` and the all-reduce(sum) of both accumulators before the division.
        disjoint = (ddp and len(self.patch_size) == data.dim() - 2 and len(slicers) == data.shape[1])
        if disjoint:
            per = -(-len(slicers) // world)
            z0, z1 = min(rank * per, len(slicers)), min((rank + 1) * per, len(slicers))
            mine = slicers[z0:z1]
        else:
            mine = slicers[rank::world]
        logits = torch.zeros((self.num_heads, *data.shape[1:]), dtype=self.results_dtype, device=dev)
        n_pred = torch.zeros(data.shape[1:], dtype=self.results_dtype, device=dev)
        gaussian = (compute_gaussian(self.patch_size, 1.0 / 8, 10, self.results_dtype, dev) if self.use_gaussian
                    else torch.ones(self.patch_size, dtype=self.results_dtype, device=dev))
        fuse = self._can_fuse(data)
        with torch.autocast(dev.type, dtype=self.autocast_dtype, enabled=dev.type == "cuda"):
            for i in range(0, len(mine), self.tile_batch):
                group = mine[i:i + self.tile_batch]
                batch = torch.stack([data[sl] for sl in group], 0)
                if fuse:
                    self._accumulate_fused(self._predict_stacked(batch), group, logits, n_pred, gaussian, data.dim() - 1)
                    continue
                pred = self._mirror_and_predict(batch).to(self.results_dtype)
                if self.use_gaussian:
                    pred *= gaussian
                for sl, p in zip(group, pred):
                    logits[sl] += p
                    n_pred[sl[1:]] += gaussian
        if disjoint:  # the one exchange step of this path: gather every rank's finished run of slices
            logits[:, z0:z1] /= n_pred[z0:z1]
            slab = torch.zeros((self.num_heads, per, *data.shape[2:]), dtype=self.results_dtype, device=dev)
            slab[:, :z1 - z0] = logits[:, z0:z1]
            parts = [torch.empty_like(slab) for _ in range(world)]
            dist.all_gather(parts, slab)
            logits = torch.cat(parts, 1)[:, :data.shape[1]]
        else:
            if ddp:  # overlapping tiles: sum the per-rank accumulators, then divide
                dist.all_reduce(logits, op=dist.ReduceOp.SUM)
                dist.all_reduce(n_pred, op=dist.ReduceOp.SUM)
            logits /= n_pred
        if bool(torch.isinf(logits).any()):
            raise RuntimeError("Encountered inf in predicted array: reduce value_scaling_factor or use fp32 results")
        return logits[(slice(None), *revert[1:])]
