"""First-contact GPU diagnostics: each stage runs in its own process under a timeout so that a hung
kernel (e.g. an mbarrier that never completes) is reported instead of stalling the whole call.

    python tools/gpu_diag.py            # run all stages
    python tools/gpu_diag.py STAGE      # run one stage in-process
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

STAGES = ["generic_fwd", "generic_bwd", "tma_fwd", "tma_bwd", "tma_z_bf16", "rl", "rl_z_16bit", "long", "cross"]


def run_stage(stage):
    import numpy as np
    import torch

    import nnuzoo_b200.selective_scan_interface as ssi
    from oracle import scan_oracle
    from tests.test_scan_gpu import _seeded

    dev = torch.device("cuda:0")

    def rel(a, b):
        a = np.asarray(a, np.float64)
        b = np.asarray(b, np.float64)
        return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))

    def scan_case(shape, force_generic, bwd, dt="float32"):
        inp, gout = _seeded(*shape, seed=1, dt=dt)
        dinp = {k: (None if v is None else v.to(dev).requires_grad_(True)) for k, v in inp.items()}
        ssi._FORCE_GENERIC = force_generic
        out, last = ssi.selective_scan_fn(dinp["u"], dinp["delta"], dinp["A"], dinp["B"], dinp["C"], dinp["D"],
                                          dinp["z"], dinp["delta_bias"], True, True)
        torch.cuda.synchronize()
        f = {k: (None if v is None else v.float()) for k, v in inp.items()}
        ro, rl = scan_oracle.selective_scan_oracle(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], f["z"],
                                                   f["delta_bias"], True, True, precision="f64")
        res = {"out": rel(out.detach().float().cpu().numpy(), ro), "last": rel(last.detach().cpu().numpy(), rl)}
        if bwd:
            out.backward(gout.to(dev))
            torch.cuda.synchronize()
            rg = scan_oracle.selective_scan_oracle_bwd(f["u"], f["delta"], f["A"], f["B"], f["C"], f["D"], f["z"],
                                                       f["delta_bias"], True, gout.float(), precision="f64")
            names = {"du": "u", "ddelta": "delta", "dA": "A", "dB": "B", "dC": "C", "dD": "D", "dz": "z",
                     "ddelta_bias": "delta_bias"}
            for g, k in names.items():
                if rg[g] is not None:
                    res[g] = rel(dinp[k].grad.float().cpu().numpy(), rg[g])
        return res

    if stage == "generic_fwd":
        print(json.dumps({"generic_fwd (2,32,4,16,600)": scan_case((2, 32, 4, 16, 600, False), True, False)}))
    elif stage == "generic_bwd":
        print(json.dumps({"generic_bwd (2,32,4,16,600)": scan_case((2, 32, 4, 16, 600, False), True, True)}))
        print(json.dumps({"generic_bwd R=1 (2,6,2,16,320,z)": scan_case((2, 6, 2, 16, 320, True), True, True)}))
    elif stage == "tma_fwd":
        print(json.dumps({"tma_fwd (2,32,4,16,1024)": scan_case((2, 32, 4, 16, 1024, False), False, False)}))
    elif stage == "tma_bwd":
        print(json.dumps({"tma_bwd (2,32,4,16,1024)": scan_case((2, 32, 4, 16, 1024, False), False, True)}))
        print(json.dumps({"tma_bwd (2,64,2,16,4096+128)": scan_case((2, 64, 2, 16, 4224, False), False, True)}))
    elif stage == "tma_z_bf16":
        print(json.dumps({"tma z fp32": scan_case((2, 16, 1, 16, 512, True), False, True)}))
        print(json.dumps({"tma z bf16": scan_case((2, 16, 1, 16, 512, True), False, True, "bfloat16")}))
        print(json.dumps({"tma z f16": scan_case((2, 16, 1, 16, 512, True), False, True, "float16")}))
    elif stage == "rl":  # row-per-lane backward: one warp per group / several warps per group (RED) / chunked
        os.environ["NZ_RL_MIN_ELTS"] = "0"
        print(json.dumps({"rl single (2,128,4,16,1024)": scan_case((2, 128, 4, 16, 1024, False), False, True)}))
        print(json.dumps({"rl red (2,128,2,16,4128)": scan_case((2, 128, 2, 16, 4128, False), False, True)}))
        os.environ["NZ_RL_ITEMS"] = "1"
        print(json.dumps({"rl one chunk (2,128,4,16,1024)": scan_case((2, 128, 4, 16, 1024, False), False, True)}))
        del os.environ["NZ_RL_ITEMS"]
    elif stage == "rl_z_16bit":
        os.environ["NZ_RL_MIN_ELTS"] = "0"
        print(json.dumps({"rl z fp32": scan_case((2, 64, 1, 16, 2048, True), False, True)}))
        print(json.dumps({"rl z bf16": scan_case((2, 64, 1, 16, 2048, True), False, True, "bfloat16")}))
        print(json.dumps({"rl f16": scan_case((2, 64, 2, 16, 2048, False), False, True, "float16")}))
    elif stage == "long":
        print(json.dumps({"cfg1 (2,768,4,16,4096)": scan_case((2, 768, 4, 16, 4096, False), False, True)}))
        print(json.dumps({"stage1 b1 (1,128,4,16,262144)": scan_case((1, 128, 4, 16, 262144, False), False, True)}))
    elif stage == "cross":
        from nnuzoo_b200 import cross_merge, cross_scan
        from oracle import cross_oracle
        x = torch.randn(2, 8, 12, 20)
        xs = cross_scan(x.to(dev))
        ok1 = bool(np.array_equal(xs.cpu().numpy(), cross_oracle.cross_scan_2d(x.numpy())))
        y = cross_merge(xs, (12, 20))
        ok2 = bool(np.array_equal(y.cpu().numpy(), cross_oracle.cross_merge_2d(xs.cpu().numpy(), 12, 20)))
        x3 = torch.randn(1, 4, 3, 5, 7)
        xs3 = cross_scan(x3.to(dev))
        ok3 = bool(np.array_equal(xs3.cpu().numpy(), cross_oracle.cross_scan_3d(x3.numpy())))
        y3 = cross_merge(xs3, (3, 5, 7))
        ok4 = bool(np.array_equal(y3.cpu().numpy(), cross_oracle.cross_merge_3d(xs3.cpu().numpy(), 3, 5, 7)))
        print(json.dumps({"cross": [ok1, ok2, ok3, ok4]}))


def main():
    if len(sys.argv) > 1:
        run_stage(sys.argv[1])
        return
    for st in STAGES:
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), st], capture_output=True, text=True,
                               timeout=240)
            tail = (r.stdout.strip() or "") + ("\nSTDERR: " + r.stderr.strip()[-1500:] if r.returncode else "")
            print(f"[{st}] rc={r.returncode} {tail}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"[{st}] TIMEOUT (hung kernel?)", flush=True)


if __name__ == "__main__":
    main()
