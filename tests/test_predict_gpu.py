"""The sliding-window accumulate kernel (nz_sw_accumulate: mirror average + gaussian multiply-accumulate, one launch per
batch of tiles) against the eager PyTorch statement of predict_from_raw_data.py:549-565, :617-623 it replaces: the
merged logits must be IDENTICAL (torch.equal) -- same roundings in the same order -- for 2-D tiles walked through a
volume (config 5's layout), overlapping 3-D tiles with 8 mirrored passes, and a plain 2-D image."""
import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _net(dims, cin, heads, seed):
    torch.manual_seed(seed)
    conv = nn.Conv2d if dims == 2 else nn.Conv3d
    return nn.Sequential(conv(cin, 8, 3, padding=1), nn.GELU(), conv(8, heads, 3, padding=1))


CASES = [
    # (image (c, *spatial), patch, step, mirror axes, tile_batch, autocast dtype, results dtype)
    ((1, 6, 64, 48), (64, 48), 0.5, (0, 1), 4, torch.bfloat16, torch.float16),     # disjoint slices, 4 passes
    ((2, 5, 40, 72), (32, 32), 0.5, (0, 1), 5, torch.float16, torch.float16),      # overlapping 2-D tiles per slice
    ((1, 24, 40, 40), (16, 32, 32), 0.5, (0, 1, 2), 3, None, torch.float32),       # 3-D tiles, 8 passes, overlap
    ((3, 70, 50), (32, 32), 0.5, (1,), 4, torch.bfloat16, torch.bfloat16),         # 2-D image, 2 passes
    ((1, 4, 32, 32), (32, 32), 0.5, (), 2, torch.float16, torch.float16),          # no mirroring
]


@pytest.mark.parametrize("case", CASES)
def test_fused_accumulate_is_bit_identical_to_eager(case):
    from nnuzoo_b200.predict import SlidingWindowPredictor
    image, patch, step, axes, tb, act, rdt = case
    dev = torch.device("cuda:0")
    heads = 3
    net = _net(len(patch), image[0], heads, 5).to(dev)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(*image, generator=g)
    kw = dict(patch_size=patch, num_heads=heads, device=dev, tile_step_size=step, use_mirroring=bool(axes),
              mirror_axes=axes or None, tile_batch=tb, autocast_dtype=act or torch.float32, results_dtype=rdt)
    import nnuzoo_b200._native as nat
    fused = SlidingWindowPredictor(net, fused_accumulate=True, **kw)
    eager = SlidingWindowPredictor(net, fused_accumulate=False, **kw)
    n0 = nat.launch_count()
    a = fused.predict_logits(x)
    assert nat.launch_count() > n0, "the fused path must launch nz_sw_accumulate"
    b = eager.predict_logits(x)
    assert a.dtype == b.dtype and a.shape == b.shape
    assert torch.equal(a, b)


def test_use_gaussian_off():
    from nnuzoo_b200.predict import SlidingWindowPredictor
    dev = torch.device("cuda:0")
    net = _net(2, 1, 2, 6).to(dev)
    x = torch.randn(1, 3, 48, 48, generator=torch.Generator().manual_seed(2))
    kw = dict(patch_size=(32, 32), num_heads=2, device=dev, use_gaussian=False, tile_batch=3,
              autocast_dtype=torch.float16, results_dtype=torch.float16)
    a = SlidingWindowPredictor(net, fused_accumulate=True, **kw).predict_logits(x)
    b = SlidingWindowPredictor(net, fused_accumulate=False, **kw).predict_logits(x)
    assert torch.equal(a, b)
