"""M2Net (nnuzoo_b200/m2net.py) against the reference network itself (tests/golden/module_m2net_64x32.npz, made by
oracle/gen_golden.py from nnunetv2/nets/m2net.py:805-971 with selective_scan_ref as the scan).

CPU leg: the module tree / glue is checked with the native ops stood in by the oracle (the product
package has no CPU path; the stand-ins are patched in by this test only).  GPU leg: the real thing, through
the C ABI.  Tolerance: north_star's fp32 1e-3 relative (max-norm per tensor) on the seven outputs.  Gradients go
through ~1700 layers with gains of 1e3 under the seeded fill: the reference's own fp32 result is 0.7 % away from
an fp64 evaluation of the same network (measured at fixture time: ours-fp64 vs reference-fp32 6.8e-3, ours-fp64
vs ours-fp32 2.5e-3), so gradient tensors are held to 2e-2 -- a glue error (wrong channel order, transposed
merge) shows up as O(1).
"""
import numpy as np
import pytest
import torch

from oracle.fill import deterministic_fill
from oracle.torch_port import selective_scan_port
from tests.helpers import load_golden, rel_err

GOLD = "module_m2net_64x32"


def _cross_scan_torch(x):
    """m2net.py:175-177 written with index maps -- differentiable torch stand-in."""
    B, D, H, W = x.shape
    row = x.reshape(B, D, H * W)
    col = x.transpose(2, 3).reshape(B, D, H * W)
    return torch.stack((row, col, row.flip(-1), col.flip(-1)), 1)


def _cross_merge_torch(out_y, spatial, mode="reference"):
    """m2net.py:202-206 + :218, same association order."""
    H, W = spatial
    B, K, D, L = out_y.shape
    t = lambda v: v.reshape(B, D, W, H).transpose(2, 3).reshape(B, D, L)
    return ((out_y[:, 0] + out_y[:, 2].flip(-1)) + t(out_y[:, 1])) + t(out_y[:, 3].flip(-1))


def _run(net, rec, device):
    x = torch.from_numpy(rec["x"]).to(device).requires_grad_(True)
    outs = net(x)
    torch.autograd.backward(list(outs), [torch.from_numpy(rec[f"gd{i}"]).to(device) for i in range(7)])
    return x, outs


def _dice_per_label(seg_pred, seg_ref, labels):
    """Dice = 2 TP / (2 TP + FP + FN) per foreground label (evaluation/evaluate_predictions.py:190-197)."""
    out = []
    for r in labels:
        mp, mr = seg_pred == r, seg_ref == r
        tp, fp, fn = int((mp & mr).sum()), int((mp & ~mr).sum()), int((~mp & mr).sum())
        out.append(float("nan") if tp + fp + fn == 0 else 2 * tp / (2 * tp + fp + fn))
    return np.array(out)


def _check_dice(outs, rec):
    """north_star's end-to-end criterion: Dice within 0.005 of the reference on identical synthetic volumes.  The
    segmentation is argmax of the full-resolution logits d0; "ground truth" is a seeded label map that agrees with the
    reference's own segmentation on ~70 % of the pixels, so the Dice values are neither 0 nor 1."""
    ref_seg = rec["d0"].argmax(1)
    our_seg = outs[0].detach().float().cpu().numpy().argmax(1)
    rng = np.random.RandomState(3)
    gt = np.where(rng.rand(*ref_seg.shape) < 0.7, ref_seg, rng.randint(0, 4, ref_seg.shape))
    d_ref, d_our = _dice_per_label(ref_seg, gt, (1, 2, 3)), _dice_per_label(our_seg, gt, (1, 2, 3))
    assert np.all(np.isfinite(d_ref)) and d_ref.min() > 0.05 and d_ref.max() < 0.95
    assert np.abs(d_ref - d_our).max() <= 0.005, (d_ref, d_our)
    assert _dice_per_label(our_seg, ref_seg, (1, 2, 3)).min() >= 0.995


def _check(net, x, outs, rec, tol, gtol=2e-2):
    _check_dice(outs, rec)
    for i, o in enumerate(outs):
        assert tuple(o.shape) == rec[f"d{i}"].shape
        assert rel_err(o.detach().cpu().numpy(), rec[f"d{i}"]) < tol, f"d{i}"
    assert rel_err(x.grad.cpu().numpy(), rec["gx"]) < gtol
    params = dict(net.named_parameters())
    names = [str(n) for n in rec["grad_norm_names"]]
    assert sorted(params) == names
    ours = np.array([0.0 if params[n].grad is None else float(params[n].grad.double().norm()) for n in names])
    ref = rec["grad_norms"]
    assert np.array_equal(ours == 0.0, ref == 0.0), "different set of parameters without gradient"
    assert np.abs(ours - ref).max() <= gtol * ref.max()
    nz = ref > 1e-3 * ref.max()
    assert (np.abs(ours - ref)[nz] / ref[nz]).max() < 5 * gtol
    for k in rec:
        if k.startswith("gp_"):
            assert rel_err(params[k[3:]].grad.cpu().numpy(), rec[k]) < gtol, k


def test_m2net_tree_matches_reference_names():
    from nnuzoo_b200.m2net import M2Net
    rec = load_golden(GOLD)
    net = M2Net(1, 4, True)
    assert sorted(n for n, _ in net.named_parameters()) == [str(n) for n in rec["grad_norm_names"]]
    assert sum(p.numel() for p in net.parameters()) == 40_953_244   # tests/golden/MANIFEST.json


def test_m2net_glue_cpu_with_oracle_ops(monkeypatch):
    import nnuzoo_b200.ss2d as ss2d
    from nnuzoo_b200.m2net import M2Net
    monkeypatch.setattr(ss2d, "cross_scan", _cross_scan_torch)
    monkeypatch.setattr(ss2d, "cross_merge", _cross_merge_torch)
    monkeypatch.setattr(ss2d, "selective_scan_fn", selective_scan_port)
    monkeypatch.setattr(ss2d, "grouped_proj", lambda x, w: torch.einsum("b k n l, k m n -> b k m l", x, w))
    rec = load_golden(GOLD)
    net = M2Net(1, 4, True).eval()
    for m in net.modules():
        if isinstance(m, ss2d.SS2D):
            m.selective_scan = selective_scan_port
    deterministic_fill(net, 7)
    x, outs = _run(net, rec, "cpu")
    _check(net, x, outs, rec, 1e-3)


@pytest.mark.gpu
def test_m2net_gpu_matches_reference():
    from nnuzoo_b200.m2net import M2Net
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    rec = load_golden(GOLD)
    net = M2Net(1, 4, True).eval()
    deterministic_fill(net, 7)
    net = net.cuda()
    x, outs = _run(net, rec, "cuda")
    _check(net, x, outs, rec, 1e-3)


@pytest.mark.gpu
def test_m2net_gpu_bf16_autocast_train_step_runs():
    """Config 2's execution mode (autocast bf16, train-mode BatchNorm / DropPath): finite loss and grads."""
    from nnuzoo_b200.m2net import get_m2net
    torch.manual_seed(0)
    net = get_m2net(1, 4, True).cuda().train()
    x = torch.randn(2, 1, 64, 64, device="cuda")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        outs = net(x)
        loss = sum(o.float().square().mean() for o in outs)
    loss.backward()
    assert torch.isfinite(loss)
    assert all(torch.isfinite(p.grad).all() for p in net.parameters() if p.grad is not None)


@pytest.mark.gpu
def test_sliding_window_on_m2net_batched_tiles_equal_one_by_one():
    """Config 5 in miniature: 2-D tiles through a (1, 3, 96, 64) volume; tile_batch 4 with stacked mirrors must give
    the logits of the reference's one-tile-at-a-time loop (tolerance of bf16 autocast and the fp16 accumulators)."""
    from nnuzoo_b200.m2net import get_m2net
    from nnuzoo_b200.predict import SlidingWindowPredictor
    torch.manual_seed(1)
    net = get_m2net(1, 4, False).cuda().eval()
    vol = torch.randn(1, 3, 96, 64)
    # bf16 autocast: a randomly initialised net with eval-mode BatchNorm overflows fp16 (the reference's default)
    kw = dict(autocast_dtype=torch.bfloat16)
    a = SlidingWindowPredictor(net, (64, 64), 4, "cuda", tile_batch=1, stack_mirrors=False, **kw).predict_logits(vol)
    b = SlidingWindowPredictor(net, (64, 64), 4, "cuda", tile_batch=4, stack_mirrors=True, **kw).predict_logits(vol)
    assert a.shape == (4, 3, 96, 64) and a.dtype == torch.float16
    assert bool(torch.isfinite(a).all())
    # a different batch size lets cuDNN / cuBLAS pick other bf16 kernels; 1700 random-init layers amplify that
    scale = float(a.float().abs().max())
    diff = (a.float() - b.float()).abs()
    assert float(diff.max()) <= 1e-1 * scale and float(diff.mean()) <= 5e-3 * scale
    assert (a.argmax(0) == b.argmax(0)).float().mean() > 0.98
