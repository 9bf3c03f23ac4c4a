"""CPU model of the CUDA kernels' decomposition (no GPU needed).

nnuzoo_b200/csrc/scan_kernels.cuh splits every row into tiles (forward: 16 lane segments x 16 steps
= 256 steps, backward: 16 x 8 = 128 steps), folds a segment sequentially, combines the lane
aggregates with a Kogge-Stone scan, hands the state from tile to tile along L (the chained {value,
tag} slots), writes a checkpoint of h every NZ_CHUNK = 128 steps (two per forward tile) and
(backward) walks the tiles last-to-first restarting from those checkpoints.  This file restates
exactly that dataflow in numpy -- same identities, same carries, same masking of the ragged tail --
and checks it against the oracle, so an algebra mistake is caught here rather than on the GPU.
"""
import numpy as np
import pytest

from oracle import scan_oracle

CK = 128                 # NZ_CHUNK: checkpoint interval
MF, LF = 16, 16          # forward: steps per lane, lane segments per row  (scan_inst.cuh NZ_FWD_M / NZ_FWD_LPR)
MB, LB = 8, 16           # backward                                      (NZ_BWD_M / NZ_BWD_LPR)


def _softplus(x):
    return np.where(x > 20, x, np.log1p(np.exp(np.minimum(x, 20))))


def _ks_inclusive(P, H):
    """Kogge-Stone inclusive scan of (P, H) pairs over the lane axis, op = later o earlier."""
    P, H = P.copy(), H.copy()
    LPR = P.shape[0]
    off = 1
    while off < LPR:
        Pp = np.concatenate([np.ones(off), P[:-off]])
        Hp = np.concatenate([np.zeros(off), H[:-off]])
        take = np.arange(LPR) >= off
        H = np.where(take, P * Hp + H, H)
        P = np.where(take, P * Pp, P)
        off *= 2
    return P, H


def _ks_suffix(Q, G):
    Q, G = Q.copy(), G.copy()
    LPR = Q.shape[0]
    off = 1
    while off < LPR:
        Qn = np.concatenate([Q[off:], np.ones(off)])
        Gn = np.concatenate([G[off:], np.zeros(off)])
        take = np.arange(LPR) + off < LPR
        G = np.where(take, Q * Gn + G, G)
        Q = np.where(take, Q * Qn, Q)
        off *= 2
    return Q, G


def model_row(u, delta, A, B, C, D, z, bias, softplus, dout):
    """One (batch, dim) row through the kernels' dataflow.  B, C: (N, L)."""
    L, N = u.shape[0], A.shape[0]
    TLF, TLB = MF * LF, MB * LB
    assert TLB == CK and TLF % CK == 0
    nchf = (L + TLF - 1) // TLF
    nck = (L + CK - 1) // CK
    pad = nchf * TLF - L  # the forward padding covers the backward's (TLB divides TLF)
    padv = lambda a: np.concatenate([a, np.zeros(pad)])  # noqa: E731  (TMA zero-fills the tail)
    valid = np.arange(nchf * TLF) < L
    uu, dr, go = padv(u), padv(delta), padv(dout)
    zz = padv(z) if z is not None else None
    Bp = np.concatenate([B, np.zeros((N, pad))], axis=1)
    Cp = np.concatenate([C, np.zeros((N, pad))], axis=1)
    x = dr + bias
    dl = np.where(valid, _softplus(x) if softplus else x, 0.0)
    dlu = dl * uu
    # ------------------------------ forward ------------------------------
    M, LPR, TL = MF, LF, TLF
    y = D * uu
    ckpt = np.zeros((nck, N))
    hc = np.zeros(N)  # the chained hand-off: h at the end of the previous tile
    for c in range(nchf):
        sl = slice(c * TL, (c + 1) * TL)
        for n in range(N):
            a = np.exp(dl[sl] * A[n]).reshape(LPR, M)
            b = (dlu[sl] * Bp[n, sl]).reshape(LPR, M)
            P, H = np.ones(LPR), np.zeros(LPR)
            for i in range(M):
                H = a[:, i] * H + b[:, i]
                P = P * a[:, i]
            Pi, Hi = _ks_inclusive(P, H)
            Pex = np.concatenate([[1.0], Pi[:-1]])
            Hex = np.concatenate([[0.0], Hi[:-1]])
            h = Pex * hc[n] + Hex
            hend = Pi * hc[n] + Hi  # state at the end of every lane segment
            for seg in range(LPR):  # lanes whose segment ends a checkpoint interval write it
                if (seg + 1) % (CK // M) == 0:
                    ck = c * (TL // CK) + (seg + 1) // (CK // M) - 1
                    if ck < nck:
                        ckpt[ck, n] = hend[seg]
            cv = Cp[n, sl].reshape(LPR, M)
            yy = y[sl].reshape(LPR, M)
            for i in range(M):
                h = a[:, i] * h + b[:, i]
                yy[:, i] += cv[:, i] * h
            y[sl] = yy.reshape(-1)
            hc[n] = hend[-1]
    out = y.copy()
    if zz is not None:
        out = out * zz / (1 + np.exp(-zz))
    # ------------------------------ backward ------------------------------
    dy = np.where(valid, go, 0.0)
    dzv = None
    if zz is not None:
        sg = 1 / (1 + np.exp(-zz))
        dzf = dy * sg * (1 + zz * (1 - sg))
        dy = dy * zz * sg
    M, LPR, TL = MB, LB, TLB
    nch = nck  # one backward tile per checkpoint interval
    tot = nchf * TLF
    du = np.zeros(tot)
    dd = np.zeros(tot)
    dA = np.zeros(N)
    dB = np.zeros((N, tot))
    dC = np.zeros((N, tot))
    yv = D * uu
    dhc = np.zeros(N)  # the chained hand-off: dh leaving the later tile
    anx = np.zeros(N)  # a of the later tile's first step (the kernel recomputes it from delta)
    for c in range(nch - 1, -1, -1):
        sl = slice(c * TL, (c + 1) * TL)
        sB = np.zeros((LPR, M))
        ddl = np.zeros((LPR, M))
        dlc, dluc, dyc = dl[sl].reshape(LPR, M), dlu[sl].reshape(LPR, M), dy[sl].reshape(LPR, M)
        for n in range(N):
            hcn = ckpt[c - 1, n] if c > 0 else 0.0
            a = np.exp(dlc * A[n])
            bv = Bp[n, sl].reshape(LPR, M)
            cv = Cp[n, sl].reshape(LPR, M)
            P, H = np.ones(LPR), np.zeros(LPR)
            for i in range(M):
                H = a[:, i] * H + dluc[:, i] * bv[:, i]
                P = P * a[:, i]
            Pi, Hi = _ks_inclusive(P, H)
            h = np.concatenate([[1.0], Pi[:-1]]) * hcn + np.concatenate([[0.0], Hi[:-1]])
            ah = np.zeros((LPR, M))
            cdy = np.zeros((LPR, M))
            for i in range(M):
                ah[:, i] = a[:, i] * h
                h = dluc[:, i] * bv[:, i] + ah[:, i]
                dC[n, sl].reshape(LPR, M)[:, i] = dyc[:, i] * h
                yv[sl].reshape(LPR, M)[:, i] += cv[:, i] * h
                cdy[:, i] = cv[:, i] * dyc[:, i]
            anl = np.concatenate([a[1:, 0], [anx[n]]])
            aup = np.concatenate([a[:, 1:], anl[:, None]], axis=1)
            Q, G = np.ones(LPR), np.zeros(LPR)
            for i in range(M - 1, -1, -1):
                G = aup[:, i] * G + cdy[:, i]
                Q = Q * aup[:, i]
            Qs, Gs = _ks_suffix(Q, G)
            dh = np.concatenate([Qs[1:], [1.0]]) * dhc[n] + np.concatenate([Gs[1:], [0.0]])
            dhnew = Qs[0] * dhc[n] + Gs[0]
            gs = np.zeros(LPR)
            for i in range(M - 1, -1, -1):
                dh = aup[:, i] * dh + cdy[:, i]
                sB[:, i] += dh * bv[:, i]
                gq = dh * ah[:, i]
                ddl[:, i] += A[n] * gq
                gs += dlc[:, i] * gq
                dB[n, sl].reshape(LPR, M)[:, i] = dh * dluc[:, i]
            dA[n] += gs.sum()
            dhc[n] = dhnew
            anx[n] = a[0, 0]
        uc = uu[sl].reshape(LPR, M)
        du[sl] = (dlc * sB + D * dyc).reshape(-1)
        gd = uc * sB + ddl
        if softplus:
            gd = gd * -np.expm1(-dlc)
        dd[sl] = gd.reshape(-1)
    dd = np.where(valid, dd, 0.0)
    if zz is not None:
        dzv = (dzf * yv)[:L]
    return dict(out=out[:L], last=ckpt[-1], du=du[:L], ddelta=dd[:L], dA=dA, dB=dB[:, :L], dC=dC[:, :L],
                dD=float((dy * uu).sum()), dbias=float(dd.sum()), dz=dzv)


@pytest.mark.parametrize("L,has_z,softplus", [(256, False, True), (700, True, True), (75, True, False),
                                              (513, False, True)])
def test_chunked_dataflow_matches_oracle(L, has_z, softplus):
    rng = np.random.default_rng(L)
    N = 16
    u = rng.standard_normal(L)
    delta = 0.5 * rng.standard_normal(L) if softplus else 0.001 + 0.1 * rng.random(L)
    A = -(np.arange(1, N + 1) * np.exp(0.1 * rng.standard_normal(N)))
    B = rng.standard_normal((N, L))
    C = rng.standard_normal((N, L))
    D, bias = 1.1, (-2.0 if softplus else 0.0)
    z = rng.standard_normal(L) if has_z else None
    go = rng.standard_normal(L)
    got = model_row(u, delta, A, B, C, D, z, bias, softplus, go)

    f = lambda a: None if a is None else np.asarray(a, np.float32)  # noqa: E731
    args = (f(u)[None, None], f(delta)[None, None], f(A)[None], f(B)[None, None], f(C)[None, None],
            f([D]), None if z is None else f(z)[None, None], f([bias]))
    out, last = scan_oracle.selective_scan_oracle(*args, delta_softplus=softplus, return_last_state=True,
                                                  precision="f64")
    g = scan_oracle.selective_scan_oracle_bwd(*args, softplus, f(go)[None, None], precision="f64")

    def close(a, b, what):
        a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
        assert np.abs(a - b).max() <= 2e-5 * max(np.abs(b).max(), 1e-12), what

    close(got["out"], out[0, 0], "out")
    close(got["last"], last[0, 0], "last_state")
    close(got["du"], g["du"][0, 0], "du")
    close(got["ddelta"], g["ddelta"][0, 0], "ddelta")
    close(got["dA"], g["dA"][0], "dA")
    close(got["dB"], g["dB"][0, 0], "dB")
    close(got["dC"], g["dC"][0, 0], "dC")
    close(got["dD"], g["dD"][0], "dD")
    close(got["dbias"], g["ddelta_bias"][0], "ddelta_bias")
    if has_z:
        close(got["dz"], g["dz"][0, 0], "dz")
