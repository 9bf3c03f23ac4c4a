"""Debug: gradients of one scan shape under different kernel selections, compared with the warp-scan kernels."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from nnuzoo_b200 import selective_scan_fn  # noqa: E402


def run(shape, env):
    for k in ("NZ_NO_RL", "NZ_RL_FWD", "NZ_RL_BWD2", "NZ_RL_NOWIDE", "NZ_RL_ITEMS"):
        os.environ.pop(k, None)
    os.environ.update(env)
    b, kd, L = shape
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    u = torch.randn(b, kd, L, device=dev, generator=g).requires_grad_(True)
    dl = (0.5 * torch.randn(b, kd, L, device=dev, generator=g)).requires_grad_(True)
    A = (-torch.arange(1, 17, device=dev).float().repeat(kd, 1)).requires_grad_(True)
    B = torch.randn(b, 4, 16, L, device=dev, generator=g).requires_grad_(True)
    C = torch.randn(b, 4, 16, L, device=dev, generator=g).requires_grad_(True)
    D = torch.ones(kd, device=dev).requires_grad_(True)
    bias = torch.full((kd,), -3.0, device=dev).requires_grad_(True)
    go = torch.randn(b, kd, L, device=dev, generator=g)
    out = selective_scan_fn(u, dl, A, B, C, D, None, bias, True)
    out.backward(go)
    torch.cuda.synchronize()
    return [out.detach()] + [t.grad for t in (u, dl, A, B, C, D, bias)]


def main():
    shape = tuple(int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (1, 128, 262144)
    ref = run(shape, {"NZ_NO_RL": "1"})
    names = ["out", "du", "ddelta", "dA", "dB", "dC", "dD", "dbias"]
    for env in ({}, {"NZ_RL_FWD": "0"}, {"NZ_RL_BWD2": "0"}, {"NZ_RL_FWD": "0", "NZ_RL_BWD2": "0"}, {"NZ_RL_NOWIDE": "1"},
                {"NZ_RL_ITEMS": "1000"}):
        got = run(shape, env)
        errs = {n: float((a - r).abs().max() / r.abs().max()) for n, a, r in zip(names, got, ref)}
        print(env, {k: f"{v:.2e}" for k, v in errs.items()})
        if "du" in errs and errs["du"] > 1e-3:
            d = (got[1] - ref[1]).abs().amax(dim=(0, 1))
            bad = torch.nonzero(d > 1e-3 * ref[1].abs().max()).flatten()
            print("   first / last bad t:", int(bad[0]), int(bad[-1]), "count", bad.numel())




def detail():
    shape = (1, 128, 262144)
    ref = run(shape, {"NZ_NO_RL": "1"})
    for rep in range(2):
        got = run(shape, {})
        d = (got[1] - ref[1]).abs()[0]  # du (rows, L)
        thr = 1e-3 * float(ref[1].abs().max())
        bad = torch.nonzero(d > thr)
        print("bad du entries:", bad.shape[0])
        rows = sorted(set(bad[:, 0].tolist()))
        print("rows:", rows[:40], "n rows", len(rows))
        ts = sorted(set(bad[:, 1].tolist()))
        print("t (first 40):", ts[:40])
        print("t mod 160 (chunk = 5 tiles):", sorted(set(t % 160 for t in ts))[:40])
        print("t // 8 blocks:", sorted(set(t // 8 for t in ts))[:40])
        dd = (got[2] - ref[2]).abs()[0]
        badd = torch.nonzero(dd > 1e-3 * float(ref[2].abs().max()))
        print("bad ddelta entries:", badd.shape[0], "rows", sorted(set(badd[:, 0].tolist()))[:20], "t//8", sorted(set((badd[:, 1] // 8).tolist()))[:20])


if len(sys.argv) > 1 and sys.argv[1] == "detail":
    detail()

if __name__ == "__main__" and not (len(sys.argv) > 1 and sys.argv[1] == "detail"):
    main()
