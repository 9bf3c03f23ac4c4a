"""numpy restatement of the SS2D / SSND CrossScan and CrossMerge permutations.

TEST INFRASTRUCTURE, NOT PRODUCT CODE (see oracle/scan_oracle.py for who may import it).

Follows, as pure data movement plus an ordered fp32 sum:
  2-D scan : nnunetv2/nets/m2net.py:175-177  (same code at SwinUMamba.py:230-232, ssnd2net.py:244-247)
  2-D merge: nnunetv2/nets/m2net.py:202-206 and the sum at :218 (ssnd2net.py:280-283)
  3-D scan : nnunetv2/nets/ssnd2net.py:250-255
  3-D merge: nnunetv2/nets/ssnd2net.py:286-298, INCLUDING the reference quirk that both the
             "wzh" and the "hwz" un-permutes read direction 1 (and its flip, direction 4) and that
             directions 2 and 5 never reach the output.

Parity status: pinned bit-exactly against the reference expressions by
oracle/gen_golden.py -> tests/golden/cross_*.npz.
"""
from __future__ import annotations

import numpy as np


def cross_scan_2d(x: np.ndarray) -> np.ndarray:
    """x (B, D, H, W) -> xs (B, 4, D, L): row-major, column-major, and their L-flips."""
    B, D, H, W = x.shape
    L = H * W
    hw = x.reshape(B, D, L)                                  # m2net.py:175 x.view(B, -1, L)
    wh = np.ascontiguousarray(x.transpose(0, 1, 3, 2)).reshape(B, D, L)  # :175 transpose(2,3)
    fwd = np.stack([hw, wh], axis=1)                         # :175-176 stack
    return np.concatenate([fwd, fwd[..., ::-1]], axis=1)     # :177 cat with flip


def cross_merge_2d(out_y: np.ndarray, H: int, W: int) -> np.ndarray:
    """out_y (B, 4, D, L) fp32 -> y (B, D, L) = ((y0 + flip(y2)) + T(y1)) + T(flip(y3))."""
    B, K, D, L = out_y.shape
    assert K == 4 and L == H * W
    inv = out_y[:, 2:4, :, ::-1]                                              # m2net.py:202
    wh = out_y[:, 1].reshape(B, D, W, H).transpose(0, 1, 3, 2).reshape(B, D, L)   # :203
    invwh = inv[:, 1].reshape(B, D, W, H).transpose(0, 1, 3, 2).reshape(B, D, L)  # :204
    y = out_y[:, 0] + inv[:, 0]                              # :218, left-to-right association
    y = y + wh
    y = y + invwh
    return np.ascontiguousarray(y.astype(np.float32))


def cross_scan_3d(x: np.ndarray) -> np.ndarray:
    """x (B, D, Z, H, W) -> xs (B, 6, D, L): orders zhw, wzh, hwz and their L-flips."""
    B, D, Z, H, W = x.shape
    L = Z * H * W
    zhw = x.reshape(B, D, L)                                                     # ssnd2net.py:250
    wzh = np.ascontiguousarray(x.transpose(0, 1, 4, 2, 3)).reshape(B, D, L)      # :251
    hwz = np.ascontiguousarray(x.transpose(0, 1, 3, 4, 2)).reshape(B, D, L)      # :252
    fwd = np.stack([zhw, wzh, hwz], axis=1)                                      # :254
    return np.concatenate([fwd, fwd[..., ::-1]], axis=1)                         # :255


def cross_merge_3d(out_y: np.ndarray, Z: int, H: int, W: int, mode: str = "reference") -> np.ndarray:
    """out_y (B, 6, D, L) -> y (B, D, L).

    mode="reference": bit-for-bit what ssnd2net.py:286-298 computes (directions 2 and 5 unused;
    the second pair re-reads direction 1 / 4 through a (W, Z, H)-shaped view whose axes are
    relabelled "h w z").  mode="fixed": the evidently intended merge (direction 2 / 5 un-permuted
    from their own hwz order); never used for parity.
    """
    B, K, D, L = out_y.shape
    assert K == 6 and L == Z * H * W
    inv = out_y[:, 3:6, :, ::-1]                                                 # :286
    v1 = out_y[:, 1].reshape(B, D, W, Z, H)
    iv1 = inv[:, 1].reshape(B, D, W, Z, H)
    y_wzh = v1.transpose(0, 1, 3, 4, 2).reshape(B, D, L)       # :291 "b c w z h -> b c z h w"
    inv_y_wzh = iv1.transpose(0, 1, 3, 4, 2).reshape(B, D, L)  # :292
    if mode == "reference":
        # :295-296 rearrange(view(B,-1,W,Z,H), "b c h w z -> b c z h w"): the three axes of the
        # (W, Z, H)-shaped view are *named* h, w, z, so the output axes are (view2, view0, view1).
        y_hwz = v1.transpose(0, 1, 4, 2, 3).reshape(B, D, L)
        inv_y_hwz = iv1.transpose(0, 1, 4, 2, 3).reshape(B, D, L)
    elif mode == "fixed":
        v2 = out_y[:, 2].reshape(B, D, H, W, Z)
        iv2 = inv[:, 2].reshape(B, D, H, W, Z)
        y_hwz = v2.transpose(0, 1, 4, 2, 3).reshape(B, D, L)
        inv_y_hwz = iv2.transpose(0, 1, 4, 2, 3).reshape(B, D, L)
    else:
        raise ValueError(mode)
    y = out_y[:, 0] + inv[:, 0]                                                  # :298
    for t in (y_wzh, inv_y_wzh, y_hwz, inv_y_hwz):
        y = y + t
    return np.ascontiguousarray(y.astype(np.float32))
