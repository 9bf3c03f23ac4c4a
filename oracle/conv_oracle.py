"""CPU restatement of the Mamba block's short convolution -- TEST INFRASTRUCTURE ONLY.

Reference: ``x = self.act(self.conv1d(x)[..., :seqlen])`` with
``nn.Conv1d(d, d, kernel_size=W, groups=d, padding=W-1)`` (nnunetv2/nets/seg_mamba/mamba_simple.py:72-82,
:316-317), i.e. out[b,d,l] = silu(bias[d] + sum_k w[d,k] * x[b,d,l-(W-1)+k]) with zeros left of the sequence.
Plain numpy in float64, plus the hand-derived adjoint; pinned to the reference expression (torch, CPU) in
tests/test_oracle_golden.py.
"""
import numpy as np


def causal_conv1d_oracle(x, w, bias=None, silu=True):
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    B, D, L = x.shape
    W = w.shape[1]
    xp = np.concatenate([np.zeros((B, D, W - 1)), x], axis=2)
    pre = np.zeros((B, D, L))
    for k in range(W):
        pre += w[None, :, k, None] * xp[:, :, k:k + L]
    if bias is not None:
        pre += np.asarray(bias, np.float64)[None, :, None]
    return pre / (1 + np.exp(-pre)) if silu else pre


def causal_conv1d_oracle_bwd(x, w, bias, dout, silu=True):
    x = np.asarray(x, np.float64)
    w = np.asarray(w, np.float64)
    dout = np.asarray(dout, np.float64)
    B, D, L = x.shape
    W = w.shape[1]
    xp = np.concatenate([np.zeros((B, D, W - 1)), x], axis=2)
    pre = np.zeros((B, D, L))
    for k in range(W):
        pre += w[None, :, k, None] * xp[:, :, k:k + L]
    if bias is not None:
        pre += np.asarray(bias, np.float64)[None, :, None]
    if silu:
        s = 1 / (1 + np.exp(-pre))
        dy = dout * s * (1 + pre * (1 - s))
    else:
        dy = dout
    dyp = np.concatenate([dy, np.zeros((B, D, W - 1))], axis=2)
    dx = np.zeros((B, D, L))
    dw = np.zeros((D, W))
    for k in range(W):
        dx += w[None, :, k, None] * dyp[:, :, W - 1 - k:W - 1 - k + L]
        dw[:, k] = (dy * xp[:, :, k:k + L]).sum(axis=(0, 2))
    return dict(dx=dx, dw=dw, dbias=dy.sum(axis=(0, 2)))
