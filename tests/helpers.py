"""Shared helpers for the parity tests (golden loading, error metrics)."""
import glob
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False) as f:
        return {k: f[k] for k in f.files}


def scan_inputs(rec):
    """(dict of input arrays or None, softplus flag, dtype name) from a scan_* golden record."""
    keys = ("u", "delta", "A", "B", "C", "D", "z", "delta_bias")
    inp = {k: rec.get("in_" + k) for k in keys}
    return inp, bool(int(rec["meta_softplus"])), str(rec["meta_dtype"])


def rel_err(a, b):
    """max |a-b| / max(|b|): the 'rel. tolerance' of BASELINE.json's north_star, per tensor."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / denom
