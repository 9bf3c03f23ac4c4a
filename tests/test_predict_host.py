"""Sliding-window predictor (nnuzoo_b200/predict.py) on CPU: tiling / gaussian against the reference's own functions
(sliding_window_prediction.py:11-56), the merge against a literal restatement of the reference loop
(predict_from_raw_data.py:549-634: batch 1, mirrors one by one), and the 2-rank sharded run (gloo) against the
single-process result.  The network is a small conv stand-in (the SS2D nets have no CPU path by design)."""
import itertools
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn as nn

from nnuzoo_b200.predict import (SlidingWindowPredictor, compute_gaussian, compute_steps_for_sliding_window,
                                 pad_to_patch, sliding_window_slicers)


def _net():
    torch.manual_seed(11)
    return nn.Sequential(nn.Conv2d(1, 6, 3, padding=1), nn.BatchNorm2d(6), nn.ReLU(), nn.Conv2d(6, 3, 1)).eval()


def _reference_loop(net, image, patch, step, mirror_axes, dtype):
    """The reference's loop, statement for statement, on one process."""
    data, revert = pad_to_patch(image, patch)
    slicers = sliding_window_slicers(data.shape[1:], patch, step)
    logits = torch.zeros((3, *data.shape[1:]), dtype=dtype)
    n_pred = torch.zeros(data.shape[1:], dtype=dtype)
    g = compute_gaussian(patch, 1.0 / 8, 10, dtype)
    combos = [c for i in range(len(mirror_axes)) for c in itertools.combinations([m + 2 for m in mirror_axes], i + 1)]
    with torch.no_grad():
        for sl in slicers:
            x = data[sl][None]
            pred = net(x)
            for c in combos:
                pred += torch.flip(net(torch.flip(x, c)), c)
            pred /= (len(combos) + 1)
            pred = pred[0].to(dtype)
            pred *= g
            logits[sl] += pred
            n_pred[sl[1:]] += g
    logits /= n_pred
    return logits[(slice(None), *revert[1:])]


def test_steps_and_slicers_known_answers():
    assert compute_steps_for_sliding_window((110,), (64,), 0.5) == [[0, 23, 46]]       # the reference's own example
    assert compute_steps_for_sliding_window((512, 512), (512, 512), 0.5) == [[0], [0]]
    sl = sliding_window_slicers((200, 512, 512), (512, 512), 0.5)
    assert len(sl) == 200 and sl[7] == (slice(None), 7, slice(0, 512), slice(0, 512))   # config 5: 200 slicers
    sl = sliding_window_slicers((96, 160, 160), (64, 128, 128), 0.5)
    assert len(sl) == 2 * 2 * 2 and sl[-1] == (slice(None), slice(32, 96), slice(32, 160), slice(32, 160))


@pytest.mark.reference
def test_steps_and_gaussian_match_reference_functions():
    pytest.importorskip("scipy")
    from oracle import ref_loader
    ref = ref_loader.load_file("ref_sliding_window", "nnunetv2/inference/sliding_window_prediction.py")
    for img, tile, step in [((110,), (64,), 0.5), ((512, 600), (512, 512), 0.5), ((200, 300, 77), (64, 128, 64), 0.3),
                            ((96, 160, 160), (96, 160, 160), 1.0), ((33, 100), (32, 48), 0.75)]:
        assert compute_steps_for_sliding_window(img, tile, step) == ref.compute_steps_for_sliding_window(img, tile, step)
    for tile in [(32, 48), (512, 512), (16, 24, 20)]:
        ours = compute_gaussian(tile, 1.0 / 8, 10, torch.float16)
        theirs = ref.compute_gaussian(tile, sigma_scale=1.0 / 8, value_scaling_factor=10, dtype=torch.float16,
                                      device=torch.device("cpu"))
        assert ours.shape == theirs.shape and float(ours.max()) == 10.0
        assert float((ours.float() - theirs.float()).abs().max()) <= 2 ** -10 * 10    # one fp16 ulp at the peak
        assert bool((ours > 0).all())


@pytest.mark.parametrize("tile_batch,stack", [(1, False), (3, True), (4, False)])
def test_predictor_equals_reference_loop(tile_batch, stack):
    net = _net()
    torch.manual_seed(2)
    image = torch.randn(1, 5, 40, 52)                 # 2-D tiles walked through 5 slices, overlapping in-plane
    want = _reference_loop(net, image, (32, 32), 0.5, (0, 1), torch.float32)
    p = SlidingWindowPredictor(net, (32, 32), 3, "cpu", tile_batch=tile_batch, stack_mirrors=stack,
                               results_dtype=torch.float32)
    got = p.predict_logits(image)
    assert got.shape == (3, 5, 40, 52)
    assert torch.allclose(got, want, atol=2e-5, rtol=1e-5)
    n_tiles = 5 * 2 * 3
    assert p.forwards == (-(-n_tiles // tile_batch)) * (1 if stack else 4)


def test_predictor_pads_small_images_and_reverts():
    net = _net()
    image = torch.randn(1, 2, 20, 30)
    p = SlidingWindowPredictor(net, (32, 32), 3, "cpu", results_dtype=torch.float32)
    got = p.predict_logits(image)
    assert got.shape == (3, 2, 20, 30)
    assert torch.allclose(got, _reference_loop(net, image, (32, 32), 0.5, (0, 1), torch.float32), atol=2e-5)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        net = _net()
        torch.manual_seed(2)
        image = torch.randn(1, 5, 40, 52)
        p = SlidingWindowPredictor(net, (32, 32), 3, "cpu", tile_batch=2, results_dtype=torch.float32)
        out = p.predict_logits(image)
        if rank == 0:
            q.put(out.numpy().copy())
    finally:
        dist.destroy_process_group()


def test_sharded_prediction_equals_single_process_gloo():
    port = 31500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    got = q.get()
    for pr in procs:
        pr.join(120)
        assert pr.exitcode == 0
    net = _net()
    torch.manual_seed(2)
    image = torch.randn(1, 5, 40, 52)
    want = SlidingWindowPredictor(net, (32, 32), 3, "cpu", tile_batch=2, results_dtype=torch.float32).predict_logits(image)
    assert np.allclose(got, want.numpy(), atol=2e-5, rtol=1e-5)
