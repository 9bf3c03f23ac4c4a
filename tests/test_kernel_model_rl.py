"""CPU model of the ROW-PER-LANE scan kernels' decomposition (nnuzoo_b200/csrc/scan_rl_kernels.cuh) -- no GPU needed.

One row (a lane) through exactly the dataflow of the round-2 kernels:
  forward : aggregate pass per chunk (P = prod a, S = state leaving the chunk from a zero start), combine walk
            h_in[c+1] = P_c h_in[c] + S_c, main pass per chunk from h_in[c] with a fine checkpoint of h every 8 steps;
  backward: aggregate pass per chunk of the reverse recurrence in R-form (dh_t = C_t dy_t + R_t, R_{t-1} = a_t dh_t):
            G = R leaving the chunk at its first step from a zero start, Q = prod a; combine walk last chunk first
            R_in(c) = G(c+1) + Q(c+1) R_in(c+1); main pass per 8-step block, last block first: h entering the block comes
            from the fine checkpoint, ONE forward recurrence for h and ONE reverse recurrence for dh per state;
  reversed groups (NzScanDesc::rev_mask): the same over t = L-1 .. 0 on the un-flipped arrays, results written at
            un-flipped positions -- must equal the oracle applied to flipped copies.
Checked against the fp64 oracle (oracle/scan_oracle.c), so an algebra mistake in chunking, carries, checkpoints or the
R-form bookkeeping is caught here rather than on the GPU.
"""
import numpy as np
import pytest

from oracle import scan_oracle

FINE = 8


def _softplus(x):
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def model_row_rl(u, delta, A, B, C, D, z, bias, softplus, dout, tiles_per_chunk, tile=32, reverse=False):
    """u, delta, z, dout: (L,); A: (N,); B, C: (N, L).  L is a whole number of `tile`-step tiles (the kernels' 128-byte
    lines).  `reverse`: walk t = L-1 .. 0 (positions stay un-flipped)."""
    L, N = u.shape[0], A.shape[0]
    assert L % tile == 0 and tile % FINE == 0
    ntl = L // tile
    nchunks = (ntl + tiles_per_chunk - 1) // tiles_per_chunk
    # walking order: index w = 0 .. L-1 of the walk visits position pos(w)
    pos = (lambda w: L - 1 - w) if reverse else (lambda w: w)
    order = np.array([pos(w) for w in range(L)])
    x = delta + bias
    dl = _softplus(x) if softplus else x
    a = np.exp(dl[None, :] * A[:, None])                      # (N, L) at positions
    bu = (dl * u)[None, :] * B                                # dl_t u_t B_t[n]
    # chunk c covers walk indices [lo_c, hi_c): whole tiles in POSITION space, walked in walking order
    bounds = []
    for c in range(nchunks):
        t_lo, t_hi = c * tiles_per_chunk, min(ntl, (c + 1) * tiles_per_chunk)
        ps = np.arange(t_lo * tile, t_hi * tile)
        bounds.append(ps[::-1] if reverse else ps)            # positions of the chunk in walking order
    chunk_walk = bounds[::-1] if reverse else bounds          # chunks in walking order

    # ---- forward: aggregate, combine, main ----
    P = np.ones((len(chunk_walk), N))
    S = np.zeros((len(chunk_walk), N))
    for ci, ps in enumerate(chunk_walk[:-1]):                 # the last chunk in walking order needs no aggregate
        h = np.zeros(N)
        for p in ps:
            h = a[:, p] * h + bu[:, p]
        S[ci] = h
        P[ci] = np.exp(A * dl[ps].sum())                      # one exponential per (chunk, state)
    h_in = np.zeros((len(chunk_walk), N))
    for ci in range(1, len(chunk_walk)):
        h_in[ci] = P[ci - 1] * h_in[ci - 1] + S[ci - 1]
    y = np.zeros(L)
    fine = {}                                                 # (chunk, block) -> h at the END of the block (walk order)
    for ci, ps in enumerate(chunk_walk):
        h = h_in[ci].copy()
        for bi in range(len(ps) // FINE):
            for p in ps[bi * FINE:(bi + 1) * FINE]:
                h = a[:, p] * h + bu[:, p]
                y[p] = (C[:, p] * h).sum()
            fine[(ci, bi)] = h.copy()
    last = fine[(len(chunk_walk) - 1, len(chunk_walk[-1]) // FINE - 1)]
    out = y + D * u
    if z is not None:
        sig = 1.0 / (1.0 + np.exp(-z))
        gate = z * sig
        dy = dout * gate
        dz = dout * out * sig * (1.0 + z * (1.0 - sig))
        out = out * gate
    else:
        dy, dz = dout, None

    # ---- backward: aggregate (R-form), combine, main ----
    cdy = C * dy[None, :]
    G = np.zeros((len(chunk_walk), N))
    Q = np.ones((len(chunk_walk), N))
    for ci, ps in enumerate(chunk_walk):
        if ci == 0:
            continue                                          # nobody needs the aggregate of the first chunk in time
        R = np.zeros(N)
        for p in ps[::-1]:
            R = a[:, p] * (cdy[:, p] + R)                     # R_{t-1} = a_t (C_t dy_t + R_t)
        G[ci] = R
        Q[ci] = np.exp(A * dl[ps].sum())
    R_in = np.zeros((len(chunk_walk), N))                     # R entering every chunk's last step
    for ci in range(len(chunk_walk) - 2, -1, -1):
        R_in[ci] = G[ci + 1] + Q[ci + 1] * R_in[ci + 1]
    du = np.zeros(L)
    dd = np.zeros(L)
    dA = np.zeros(N)
    dB = np.zeros((N, L))
    dC = np.zeros((N, L))
    for ci, ps in enumerate(chunk_walk):
        R = R_in[ci].copy()
        nb = len(ps) // FINE
        for bi in range(nb - 1, -1, -1):
            blk = ps[bi * FINE:(bi + 1) * FINE]
            h0 = fine[(ci, bi - 1)] if bi > 0 else h_in[ci]   # h entering the block: the fine checkpoint before it
            hs, hp = [], h0.copy()
            for p in blk:                                     # ONE forward recurrence
                hprev = hp
                hp = a[:, p] * hp + bu[:, p]
                hs.append((hprev, hp))
            for j in range(FINE - 1, -1, -1):                 # ONE reverse recurrence
                p = blk[j]
                hprev, hcur = hs[j]
                dh = cdy[:, p] + R
                dC[:, p] = dy[p] * hcur
                dB[:, p] = dh * dl[p] * u[p]
                du[p] = dy[p] * D + dl[p] * (dh * B[:, p]).sum()
                gq = dh * hprev * a[:, p]
                dd[p] = (dh * B[:, p]).sum() * u[p] + (A * gq).sum()
                dA += gq * dl[p]
                R = a[:, p] * dh
    if softplus:
        dd = dd * (1.0 / (1.0 + np.exp(-x)))
    return dict(out=out, last=last, du=du, ddelta=dd, dA=dA, dB=dB, dC=dC, dD=float((dy * u).sum()),
                dbias=float(dd.sum()), dz=dz, fine=fine, nchunks=len(chunk_walk))


def _oracle(u, delta, A, B, C, D, z, bias, softplus, go):
    f = lambda v: None if v is None else np.asarray(v, np.float32)  # noqa: E731
    args = (f(u)[None, None], f(delta)[None, None], f(A)[None], f(B)[None, None], f(C)[None, None],
            f([D]), None if z is None else f(z)[None, None], f([bias]))
    out, last = scan_oracle.selective_scan_oracle(*args, delta_softplus=softplus, return_last_state=True, precision="f64")
    g = scan_oracle.selective_scan_oracle_bwd(*args, softplus, f(go)[None, None], precision="f64")
    return out[0, 0], last[0, 0], g


def _inputs(L, has_z, softplus, seed):
    rng = np.random.default_rng(seed)
    N = 16
    # (values rounded to fp32 first: the oracle takes fp32 inputs and computes in fp64)
    r32 = lambda v: np.asarray(v, np.float32).astype(np.float64)  # noqa: E731
    u = r32(rng.standard_normal(L))
    delta = r32(0.5 * rng.standard_normal(L) if softplus else 0.001 + 0.1 * rng.random(L))
    A = r32(-(np.arange(1, N + 1) * np.exp(0.1 * rng.standard_normal(N))))
    B, C = r32(rng.standard_normal((N, L))), r32(rng.standard_normal((N, L)))
    z = r32(rng.standard_normal(L)) if has_z else None
    go = r32(rng.standard_normal(L))
    return u, delta, A, B, C, float(np.float32(1.1)), z, float(np.float32(-2.0 if softplus else 0.0)), go


def _close(a, b, what, tol=2e-5):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    assert np.abs(a - b).max() <= tol * max(np.abs(b).max(), 1e-12), what


@pytest.mark.parametrize("L,tpc,has_z,softplus", [(256, 8, False, True), (256, 3, False, True), (992, 5, True, True),
                                                  (64, 1, True, False), (32, 4, False, True)])
def test_row_per_lane_dataflow_matches_oracle(L, tpc, has_z, softplus):
    u, delta, A, B, C, D, z, bias, go = _inputs(L, has_z, softplus, seed=L + tpc)
    got = model_row_rl(u, delta, A, B, C, D, z, bias, softplus, go, tiles_per_chunk=tpc)
    out, last, g = _oracle(u, delta, A, B, C, D, z, bias, softplus, go)
    _close(got["out"], out, "out")
    _close(got["last"], last, "last_state")
    for k, ref in (("du", g["du"][0, 0]), ("ddelta", g["ddelta"][0, 0]), ("dA", g["dA"][0]), ("dB", g["dB"][0, 0]),
                   ("dC", g["dC"][0, 0]), ("dD", g["dD"][0]), ("dbias", g["ddelta_bias"][0])):
        _close(got[k], ref, k)
    if has_z:
        _close(got["dz"], g["dz"][0, 0], "dz")


@pytest.mark.parametrize("tpc", [1, 3, 100])
def test_chunk_count_does_not_change_the_model(tpc):
    """The decomposition is exact: one chunk (no aggregate pass), a few, one tile per chunk."""
    u, delta, A, B, C, D, z, bias, go = _inputs(320, True, True, seed=3)
    ref = model_row_rl(u, delta, A, B, C, D, z, bias, True, go, tiles_per_chunk=100)
    got = model_row_rl(u, delta, A, B, C, D, z, bias, True, go, tiles_per_chunk=tpc)
    for k in ("out", "du", "ddelta", "dA", "dB", "dC"):
        _close(got[k], ref[k], k, tol=1e-10)


@pytest.mark.parametrize("L,tpc", [(256, 3), (96, 1)])
def test_reversed_walk_equals_the_oracle_on_flipped_copies(L, tpc):
    """rev_mask: the recurrence from t = L-1 down to 0 over the un-flipped arrays, results at un-flipped positions, equals
    flip(oracle(flip(inputs))) -- SS2D's flipped directions without flipped copies (m2net.py:176, :202-206)."""
    u, delta, A, B, C, D, z, bias, go = _inputs(L, False, True, seed=11 + L)
    got = model_row_rl(u, delta, A, B, C, D, None, bias, True, go, tiles_per_chunk=tpc, reverse=True)
    fl = lambda v: np.ascontiguousarray(v[..., ::-1])  # noqa: E731
    out, last, g = _oracle(fl(u), fl(delta), A, fl(B), fl(C), D, None, bias, True, fl(go))
    _close(got["out"], out[::-1], "out")
    _close(got["last"], last, "last_state (describes t = 0 of the un-flipped sequence)")
    _close(got["du"], g["du"][0, 0][::-1], "du")
    _close(got["ddelta"], g["ddelta"][0, 0][::-1], "ddelta")
    _close(got["dA"], g["dA"][0], "dA")
    _close(got["dB"], g["dB"][0, 0][:, ::-1], "dB")
    _close(got["dC"], g["dC"][0, 0][:, ::-1], "dC")
