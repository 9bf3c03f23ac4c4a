"""Decomposition algebra of the round-2 kernels, restated in numpy and checked against the CPU oracles -- no GPU.

  * tiled 3-D CrossScan / CrossMerge (csrc/cross_kernels.cu): every 3-D direction is the column-major walk
    t(p) = w * H + h of a RE-FACTORED (Hm x Wm) matrix, so the 2-D tile kernels serve it; the merge runs as two passes
    that keep the reference's left-to-right association, the merge adjoint accumulates direction 1 / 4 twice;
  * anti-causal convolution (csrc/conv1d_kernels.cu, REV): in a thread's mirrored step index the flipped-sequence
    convolution is the causal one, with the window loaded from the other side and reversed;
  * the row-per-lane backward's dB / dC butterfly (csrc/scan_rl_kernels.cuh): 5 select-free rounds in which every lane
    keeps half of its slots and adds the partner's, ending with each lane owning whole-warp sums;
  * the fused sliding-window accumulate (csrc/sw_kernels.cu): the mirrored passes read back through their own flips.
"""
import itertools

import numpy as np
import pytest

from oracle import conv_oracle, cross_oracle


# ---------------------------------------------------------------------------------------------------------------
def _pass_scan(x, Hm, Wm):
    """One tiled scan pass on rows viewed as an (Hm, Wm) matrix: (copy, flip of copy, column-major walk, its flip)."""
    rows = x.reshape(*x.shape[:2], Hm, Wm)
    copy = rows.reshape(*x.shape[:2], -1)
    tr = rows.transpose(0, 1, 3, 2).reshape(*x.shape[:2], -1)
    return copy, copy[..., ::-1], tr, tr[..., ::-1]


@pytest.mark.parametrize("shape", [(2, 3, 4, 5, 6), (1, 2, 3, 7, 2), (1, 1, 5, 1, 4)])
def test_3d_directions_are_column_major_walks_of_refactored_matrices(shape):
    rng = np.random.default_rng(0)
    x = rng.standard_normal(shape).astype(np.float32)
    B, D, Z, H, W = shape
    want = cross_oracle.cross_scan_3d(x)                       # (B, 6, D, L): zhw, wzh, hwz and their flips
    xf = x.reshape(B, D, -1)
    k0, k3, k1, k4 = _pass_scan(xf, Z * H, W)                  # pass A: slots 0, 3, 1, 4
    _, _, k2, k5 = _pass_scan(xf, Z, H * W)                    # pass B: slots 2, 5
    got = np.stack([k0, k1, k2, k3, k4, k5], 1)
    assert np.array_equal(got, want)


def _tr_read(a, Hm, Wm):
    """value read at output position p = h * Wm + w from the column-major walk stored in a: a[w * Hm + h]."""
    return a.reshape(*a.shape[:-1], Wm, Hm).swapaxes(-1, -2).reshape(*a.shape[:-1], -1)


@pytest.mark.parametrize("shape", [(2, 3, 4, 5, 6), (1, 2, 3, 7, 2)])
@pytest.mark.parametrize("mode", ["reference", "fixed"])
def test_3d_merge_as_two_passes_keeps_the_association_order(shape, mode):
    rng = np.random.default_rng(1)
    B, D, Z, H, W = shape
    L = Z * H * W
    oy = rng.standard_normal((B, 6, D, L)).astype(np.float32)
    want = cross_oracle.cross_merge_3d(oy, Z, H, W, mode=mode)
    # pass 1 on the (Z*H) x W matrix: ((y0 + inv0) + wzh) + inv_wzh
    acc = oy[:, 0] + oy[:, 3][..., ::-1]
    acc = acc + _tr_read(oy[:, 1], Z * H, W)
    acc = acc + _tr_read(oy[:, 4][..., ::-1], Z * H, W)
    # pass 2 adds its two terms onto y: reference mode reads directions 1 / 4 again through the H x (W*Z) matrix
    if mode == "reference":
        acc = acc + _tr_read(oy[:, 1], H, W * Z)
        acc = acc + _tr_read(oy[:, 4][..., ::-1], H, W * Z)
    else:
        acc = acc + _tr_read(oy[:, 2], Z, H * W)
        acc = acc + _tr_read(oy[:, 5][..., ::-1], Z, H * W)
    assert np.array_equal(acc.astype(np.float32), want)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("L,W,steps", [(37, 4, 8), (64, 3, 16), (5, 2, 8), (16, 4, 8)])
def test_mirrored_window_turns_the_flipped_convolution_into_the_causal_one(L, W, steps):
    """conv1d_fwd_kernel<T, REV>: a thread owning steps [l0, l0 + STEPS) loads x[l0 .. l0 + STEPS + 4) (zero outside),
    reverses it, runs the CAUSAL tap loop, reverses the results."""
    rng = np.random.default_rng(2)
    x = rng.standard_normal((1, 2, L)).astype(np.float32)
    w = rng.standard_normal((2, W)).astype(np.float32)
    b = rng.standard_normal(2).astype(np.float32)
    want = conv_oracle.causal_conv1d_oracle(np.ascontiguousarray(x[..., ::-1]), w, b, False)[..., ::-1]
    got = np.zeros_like(x)
    WMAX = 4
    sh = WMAX - W
    for d in range(2):
        for l0 in range(0, L, steps):
            win = np.array([x[0, d, p] if 0 <= p < L else 0.0 for p in range(l0, l0 + steps + 4)], np.float32)
            v = win[::-1]                                     # v[j] = x[l0 + STEPS + 3 - j]
            o = np.zeros(steps, np.float32)
            for i in range(steps):
                p = b[d]
                for k in range(WMAX):
                    if k - sh >= 0:
                        p = p + w[d, k - sh] * v[1 + i + k]
                o[i] = p
            o = o[::-1]
            for i in range(steps):
                if l0 + i < L:
                    got[0, d, l0 + i] = o[i]
    assert np.allclose(got, want, rtol=1e-5, atol=1e-6)


# ---------------------------------------------------------------------------------------------------------------
def test_select_free_butterfly_reduces_every_slot_over_the_warp():
    """32 lanes x 32 slots -> every lane ends with ONE slot summed over all lanes, in 5 rounds; in round r a lane keeps
    the half of its remaining slots selected by bit r of its lane id and adds the partner lane's copy of that half."""
    rng = np.random.default_rng(3)
    v = rng.standard_normal((32, 32))                          # [lane][slot]
    live = {lane: list(range(32)) for lane in range(32)}       # slots each lane still carries
    val = {lane: {s: v[lane, s] for s in range(32)} for lane in range(32)}
    for r in range(5):
        bit = 1 << r
        nxt_val, nxt_live = {}, {}
        for lane in range(32):
            partner = lane ^ bit
            mine = live[lane]
            half = len(mine) // 2
            keep = mine[half:] if lane & bit else mine[:half]
            # the partner keeps the other half and sends the half this lane keeps
            nxt_val[lane] = {s: val[lane][s] + val[partner][s] for s in keep}
            nxt_live[lane] = keep
        val, live = nxt_val, nxt_live
    owners = {}
    for lane in range(32):
        assert len(live[lane]) == 1
        s = live[lane][0]
        owners[s] = lane
        assert np.isclose(val[lane][s], v[:, s].sum())
    assert sorted(owners) == list(range(32))                   # every slot has exactly one owner lane


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("tile", [(1, 4, 6), (3, 4, 2)])
def test_mirrored_passes_read_back_through_their_own_flip(tile):
    """sw_accumulate_kernel: pass m was computed on the tile flipped along the axes of mask m; element e of the average
    reads pass m at the flipped position (predict_from_raw_data.py:549-565)."""
    rng = np.random.default_rng(4)
    axes = [a for a in range(3) if tile[a] > 1]
    combos = [c for i in range(len(axes)) for c in itertools.combinations(axes, i + 1)]
    f = lambda t: np.tanh(t) + 0.1 * np.arange(t.size).reshape(t.shape)   # a position-dependent "network"  # noqa: E731
    x = rng.standard_normal(tile)
    want = f(x)
    for c in combos:
        want = want + np.flip(f(np.flip(x, c)), c)
    want = want / (len(combos) + 1)
    passes = [f(x)] + [f(np.flip(x, c)) for c in combos]
    masks = [0] + [sum(1 << a for a in c) for c in combos]
    got = np.zeros(tile)
    for e in itertools.product(*[range(n) for n in tile]):
        acc = passes[0][e]
        for m in range(1, len(passes)):
            fe = tuple(tile[a] - 1 - e[a] if (masks[m] >> a) & 1 else e[a] for a in range(3))
            acc = acc + passes[m][fe]
        got[e] = acc / len(passes)
    assert np.allclose(got, want)


# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("pd,rd", [("float16", "float16"), ("bfloat16", "float16"), ("bfloat16", "bfloat16"),
                                   ("float32", "float32"), ("float16", "float32")])
def test_rounding_sequence_of_the_accumulate_equals_the_eager_expressions(pd, rd):
    """sw_accumulate_kernel computes in fp32 and rounds to the prediction dtype after every add and after the division,
    to the accumulator dtype after the conversion, the gaussian multiply and each accumulate.  That sequence, written out
    with explicit roundings, must equal the reference's in-place 16-bit tensor expressions
    (predict_from_raw_data.py:549-565, :617-623) bit for bit."""
    import torch
    pdt, rdt = getattr(torch, pd), getattr(torch, rd)
    g = torch.Generator().manual_seed(0)
    passes = [torch.randn(3, 5, 7, generator=g).to(pdt) for _ in range(4)]          # already flipped back
    gauss = (torch.rand(5, 7, generator=g) * 10).to(rdt)
    logits0 = torch.randn(3, 5, 7, generator=g).to(rdt)
    n0 = torch.rand(5, 7, generator=g).to(rdt)
    # eager (the reference's statement)
    pred = passes[0].clone()
    for p in passes[1:]:
        pred += p
    pred /= len(passes)
    pred = pred.to(rdt)
    pred *= gauss
    logits = logits0.clone()
    logits += pred
    n = n0.clone()
    n += gauss
    # the kernel's sequence: fp32 arithmetic, one rounding per step
    rp = lambda t: t.to(pdt).float()  # noqa: E731
    rr = lambda t: t.to(rdt).float()  # noqa: E731
    acc = passes[0].float()
    for p in passes[1:]:
        acc = rp(acc + p.float())
    acc = rp(acc * (1.0 / len(passes)))
    v = rr(acc)
    v = rr(v * gauss.float())
    lk = rr(logits0.float() + v).to(rdt)
    nk = rr(n0.float() + gauss.float()).to(rdt)
    assert torch.equal(lk, logits) and torch.equal(nk, n)
