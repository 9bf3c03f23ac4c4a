#!/usr/bin/env python
"""bench.py -- the hot path of BASELINE.json measured on B200.

One "step" = one pass of the hot path over one batch: the 80 SS2D selective scans (forward AND
backward, fp32 scan I/O as the reference's SS2D forces at m2net.py:185-191) of one SS2D2Net/M2Net
training step at BASELINE configs[1]: batch 12 x 1x512x512 (scan list: SURVEY.md section 8a).

  value      whole-job GB/s, algorithmic bytes of the selective_scan_fn boundary, 4*(8E + 6S) per
             fwd+bwd scan, inputs resident in HBM; every scan reads a fresh slice of
             >1 GB operand buffers, so nothing is L2-warm ("inputs larger than L2").
  e2e        the same metric through the host-buffer C-ABI call nz_scan_fwd_bwd_host: pinned host
             operands in, all results back to the host, copies inside the timed region.
  roofline   all launches of the dominant kernel (scan_bwd) inside the timed region, CUDA events on
             the launching stream, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline / --impl reference
             the reference's CPU path (selective_scan_ref restated in oracle/torch_port.py; the
             reference itself is not on the GPU box) on a bounded sample of the same workload.

  train      the second half of BASELINE.json's metric: full SS2D2Net (M2Net) optimisation steps through
             nnuzoo_b200.train.Trainer (pinned-host batch in, bf16 autocast, Dice+CE deep supervision, backward + DDP
             all-reduce over NCCL, clip, AdamW, loss read back), 12 patches of 1x512x512 per GPU: patches/s.
  infer      BASELINE configs[4]: nnuzoo_b200.predict.SlidingWindowPredictor over a synthetic 1x200x512x512 volume,
             slices sharded over ranks in contiguous runs, gaussian fp16 accumulators, one all-gather: volumes/s, slices/s.

Multi-GPU (torchrun, one rank per GPU): the scan has no cross-row dependency and the path has no
exchange step, so every rank runs the full per-GPU batch (weak scaling, "replicas only", no
collective on the data path); time is the max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_STATE = 16
K_DIR = 4
BATCH = 12


def m2net_scan_list(res: int = 512):
    """(K*D, L) of every selective_scan_fn call of one M2Net forward (m2net.py:810-872; MU stages
    :646-655, :405-427, :461-474): stage s has mid channels 16*2^s (K*D = 8*mid), depth n = 7-s and
    input resolution res/2^s; encoder scans L = r^2, (r/2)^2, ..., last one twice; decoder the same
    without the duplicate; each stage appears twice (encoder side and decoder side of the U)."""
    scans = []
    for s in range(4):
        kd = 8 * 16 * 2 ** s
        n = 7 - s
        r = res >> s
        enc = [(r >> i) ** 2 for i in range(n - 1)]
        enc.append(enc[-1])
        dec = enc[:-1]
        for _ in range(2):
            scans += [(kd, L) for L in enc + dec]
    return scans


def scan_bytes(batch, kd, L, w=4):
    E = batch * kd * L
    S = batch * K_DIR * N_STATE * L
    return dict(E=E, S=S, fwd=w * (3 * E + 2 * S), bwd=w * (5 * E + 4 * S), total=w * (8 * E + 6 * S))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].startswith("Active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def traffic_per_launch():
    """Average DRAM bytes (read + write) per nz_scan_bwd call of this workload (all the launches of the call), from the
    committed ncu capture of the same command (profiles/r02_bwd_traffic.json, written by tools/ncu_traffic.py, stamped
    with the git revision it was taken at)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r02_bwd_traffic.json")) as f:
            j = json.load(f)
            return float(j["dram_bytes_per_launch"]), j.get("git") if not j.get("note") else f'{j.get("git")} ({j["note"]})'
    except Exception:
        return None, None


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# -------------------------------------------------------------------------------------------------
# CPU arm: the reference's own CPU path on a bounded sample of the workload
# -------------------------------------------------------------------------------------------------
CFG0 = dict(batch=2, K=4, D=192, N=16, L=64 * 64, R=6)  # BASELINE.json configs[0]; R = dt_rank of d_model 96


def cfg0_tensors(device="cpu", seed=0):
    """BASELINE.md section 4 / SURVEY.md 8(d) inputs of configs[0]: B, C are the split views SS2D hands over."""
    import math

    import torch
    c = CFG0
    kd = c["K"] * c["D"]
    g = torch.Generator().manual_seed(seed)
    u = torch.randn(c["batch"], kd, c["L"], generator=g)
    delta = 0.5 * torch.randn(c["batch"], kd, c["L"], generator=g)
    A = -torch.arange(1, c["N"] + 1).float().repeat(kd, 1).contiguous()
    xdbl = torch.randn(c["batch"], c["K"], c["R"] + 2 * c["N"], c["L"], generator=g)
    dt = torch.exp(torch.rand(kd, generator=g) * (math.log(0.1) - math.log(0.001)) + math.log(0.001))
    bias = dt + torch.log(-torch.expm1(-dt))          # m2net.py:128-135
    D = torch.ones(kd)                                # m2net.py:161
    xdbl = xdbl.to(device)
    Bv, Cv = xdbl[:, :, c["R"]:c["R"] + c["N"]], xdbl[:, :, c["R"] + c["N"]:]
    return [t.to(device) for t in (u, delta, A)] + [Bv, Cv] + [t.to(device) for t in (D, bias)]


def cpu_sample_run(steps: int, warmup: int):
    """the reference's CPU path on configs[0] (B=2, K=4, D=192, N=16, L=4096, fp32, z=None,
    delta_softplus, delta_bias), FORWARD ONLY, all host threads.  /root/reference is not on the GPU box, so the
    function timed is the restatement oracle/torch_port.py of selective_scan_ref (same Python loop of torch ops over L,
    pinned to the verbatim reference by tests/golden): kind "port", labelled "restated"."""
    import platform

    import torch

    from oracle.torch_port import selective_scan_port

    torch.set_num_threads(os.cpu_count() or 1)
    u, delta, A, Bv, Cv, D, bias = cfg0_tensors("cpu")
    times = []
    with torch.no_grad():
        for _ in range(warmup + steps):
            t0 = time.perf_counter()
            out = selective_scan_port(u, delta, A, Bv, Cv, D, None, bias, True)
            times.append(time.perf_counter() - t0)
    assert bool(torch.isfinite(out).all())
    ts = sorted(times[warmup:])
    t = ts[len(ts) // 2]
    c = CFG0
    nbytes = scan_bytes(c["batch"], c["K"] * c["D"], c["L"])["fwd"]
    cpu_model = platform.processor() or platform.machine()
    try:
        with open("/proc/cpuinfo") as f:
            cpu_model = next(l.split(":", 1)[1].strip() for l in f if l.startswith("model name"))
    except Exception:
        pass
    return dict(seconds=t, gbps=nbytes / t / 1e9, cores=torch.get_num_threads(),
                sample=f"configs[0] forward only (B=2, K=4, D=192, N=16, L=4096, fp32, split-view B/C): selective_scan_ref "
                       f"restated in oracle/torch_port.py, {torch.get_num_threads()} threads on {cpu_model}; median of "
                       f"{steps} after {warmup} warm-up; {nbytes / 1e6:.1f} MB algorithmic (4(3E+2S))")


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    r = cpu_sample_run(max(1, min(args.steps, 3)), 1)  # one warm-up pass; each pass takes seconds
    line = {
        "impl": "reference", "metric": METRIC, "value": r["gbps"], "unit": "GB/s", "n_gpus": args.gpus,
        "steps": max(1, min(args.steps, 3)), "warmup": 1, "ms_per_step": r["seconds"] * 1e3,
        "note": "CPU arm as BASELINE.md section 4 fixes it: configs[0], forward only, restated selective_scan_ref; the GPU "
                "arm's line carries the same shape under configs.cfg0 (fwd and fwd+bwd) next to the configs[1] headline",
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": CONFIG,
        "cpu_baseline": {"value": r["gbps"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["gbps"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


METRIC = "SS2D selective-scan fwd+bwd GB/s (selective_scan_fn boundary bytes 4*(8E+6S))"
CONFIG = {"workload": "configs[1]: all 80 SS2D selective scans (fwd+bwd) of one SS2D2Net/M2Net training step, "
                      "batch 12 x 1x512x512, fp32 scan I/O, d_state 16, K=4",
          "scans_per_step": 80, "batch": BATCH, "l2": "inputs larger than L2 (each scan reads a fresh slice of "
                                                      ">1 GB buffers)",
          "parallelism": "replicas (one full batch per GPU, no data-path collective)"}


def max_over_ranks(value: float, dev, world: int) -> float:
    """Multi-GPU numbers are the max over ranks (NCCL on the GPU box, gloo in the CPU tests)."""
    if world <= 1:
        return value
    import torch
    import torch.distributed as dist
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# -------------------------------------------------------------------------------------------------
# GPU arm
# -------------------------------------------------------------------------------------------------
class Workload:
    def __init__(self, dev):
        import torch

        from nnuzoo_b200 import _native
        from nnuzoo_b200._native import NzScanDesc

        self.torch, self.native, self.Desc = torch, _native, NzScanDesc
        self.dev = dev
        self.lib = _native.lib()
        _native.bind_device(dev.index)
        self.scans = m2net_scan_list(512)
        kd_max_elems = max(BATCH * kd * L for kd, L in self.scans)
        bc_max = max(BATCH * K_DIR * N_STATE * L for _, L in self.scans)
        self.pool_mult = 2  # operand pools are 2x the largest scan so slices rotate
        g = torch.Generator(device=dev).manual_seed(1234 + dev.index)
        rnd = lambda n, scale=1.0: torch.randn(n, device=dev, generator=g) * scale  # noqa: E731
        self.out = torch.empty(kd_max_elems, device=dev)
        self.du = torch.empty(kd_max_elems, device=dev)
        self.dd = torch.empty(kd_max_elems, device=dev)
        self.dB = torch.empty(bc_max, device=dev)
        self.dC = torch.empty(bc_max, device=dev)
        kdm = max(kd for kd, _ in self.scans)
        self.A = -torch.arange(1, N_STATE + 1, device=dev).float().repeat(kdm, 1).contiguous()
        self.D = torch.ones(kdm, device=dev)
        self.bias = torch.full((kdm,), -2.0, device=dev)
        self.dA = torch.zeros(kdm, N_STATE, device=dev)
        self.dD = torch.zeros(kdm, device=dev)
        self.dbias = torch.zeros(kdm, device=dev)
        ck = self.native.NZ_CHUNK
        nch_max = max(BATCH * kd * ((L + ck - 1) // ck) * N_STATE for kd, L in self.scans)
        self.x = torch.empty(nch_max, device=dev)
        # fine checkpoints (h every 8 steps) for the scans the row-per-lane backward takes, and the scratch both
        # directions need (tile tickets / carries, chunk aggregates): sized once for the largest scan
        fine_max, ws_max = 0, self.native.workspace_bytes(BATCH, kdm)
        self.u = self.delta = self.Bm = self.Cm = torch.empty(64, device=dev)  # aligned stand-ins for the size queries
        for kd, L in self.scans:
            d = self._bare_desc(kd, L)
            nfine = int(self.lib.nz_scan_fine_bytes(ctypes.byref(d)))
            fine_max = max(fine_max, nfine)
            if nfine:
                d.xf = ctypes.c_void_p(256)
            ws_max = max(ws_max, int(self.lib.nz_scan_workspace_bytes_bwd(ctypes.byref(d))),
                         int(self.lib.nz_scan_workspace_bytes_cp(ctypes.byref(d))))
        self.xf = torch.empty(max(fine_max // 4, 1), device=dev)
        self.ws = torch.empty(ws_max, dtype=torch.uint8, device=dev)
        self.ws_bytes = ws_max
        self.u = rnd(kd_max_elems * self.pool_mult)
        self.delta = rnd(kd_max_elems * self.pool_mult, 0.5)
        self.dout = rnd(kd_max_elems * self.pool_mult)
        self.Bm = rnd(bc_max * self.pool_mult)
        self.Cm = rnd(bc_max * self.pool_mult)
        self.cur_row = 0
        self.cur_bc = 0
        self.total_bytes = sum(scan_bytes(BATCH, kd, L)["total"] for kd, L in self.scans)
        self.bwd_bytes = sum(scan_bytes(BATCH, kd, L)["bwd"] for kd, L in self.scans)

    def _bare_desc(self, kd, L):
        d = self.Desc()
        d.batch, d.dim, d.dstate, d.ngroups, d.seqlen = BATCH, kd, N_STATE, K_DIR, L
        d.dtype, d.delta_softplus = 0, 1
        p = lambda buf: ctypes.c_void_p(buf.data_ptr())  # noqa: E731
        d.u, d.delta, d.B, d.C = p(self.u), p(self.delta), p(self.Bm), p(self.Cm)
        for s in (d.u_stride, d.delta_stride, d.out_stride, d.dout_stride):
            s[0], s[1] = kd * L, L
        for s in (d.B_stride, d.C_stride):
            s[0], s[1], s[2] = K_DIR * N_STATE * L, N_STATE * L, L
        d.A_stride = N_STATE
        return d

    def desc_for(self, kd, L):
        t = self.torch
        n_row = BATCH * kd * L
        n_bc = BATCH * K_DIR * N_STATE * L
        if self.cur_row + n_row > self.u.numel():
            self.cur_row = 0
        if self.cur_bc + n_bc > self.Bm.numel():
            self.cur_bc = 0
        r0, b0 = self.cur_row, self.cur_bc
        self.cur_row += n_row
        self.cur_bc += n_bc
        d = self.Desc()
        d.batch, d.dim, d.dstate, d.ngroups, d.seqlen = BATCH, kd, N_STATE, K_DIR, L
        d.dtype, d.delta_softplus = 0, 1
        p = lambda buf, off=0: ctypes.c_void_p(buf.data_ptr() + 4 * off)  # noqa: E731
        d.u, d.delta, d.dout = p(self.u, r0), p(self.delta, r0), p(self.dout, r0)
        d.B, d.C = p(self.Bm, b0), p(self.Cm, b0)
        d.A, d.D, d.delta_bias = p(self.A), p(self.D), p(self.bias)
        for s in (d.u_stride, d.delta_stride, d.out_stride, d.dout_stride):
            s[0], s[1] = kd * L, L
        for s in (d.B_stride, d.C_stride):
            s[0], s[1], s[2] = K_DIR * N_STATE * L, N_STATE * L, L
        d.A_stride = N_STATE
        d.out, d.x = p(self.out), p(self.x)
        d.workspace, d.workspace_bytes = p(self.ws), self.ws_bytes
        d.du, d.ddelta = p(self.du), p(self.dd)
        d.dA, d.dB, d.dC, d.dD, d.ddelta_bias = p(self.dA), p(self.dB), p(self.dC), p(self.dD), p(self.dbias)
        if self.lib.nz_scan_fine_bytes(ctypes.byref(d)):  # this scan qualifies for the row-per-lane backward
            d.xf = p(self.xf)
        _ = t
        return d, n_bc

    def step(self, stream, bwd_events=None):
        """All 80 scans forward then backward (kernel launches + the dB/dC zero fills they need)."""
        torch = self.torch
        sp = ctypes.c_void_p(stream.cuda_stream)
        for kd, L in self.scans:
            d, n_bc = self.desc_for(kd, L)
            self.native.check(self.lib.nz_scan_fwd(ctypes.byref(d), sp), "nz_scan_fwd")
            if not self.lib.nz_scan_bwd_overwrites_dbc(ctypes.byref(d)):  # several tiles share a dB/dC element
                self.dB[:n_bc].zero_()
                self.dC[:n_bc].zero_()
            if bwd_events is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
            self.native.check(self.lib.nz_scan_bwd(ctypes.byref(d), sp), "nz_scan_bwd")
            if bwd_events is not None:
                e1.record(stream)
                bwd_events.append((e0, e1))


def e2e_run(dev, stream, steps, warmup):
    """The same workload through the host-buffer C-ABI entry point (pinned host operands)."""
    import torch

    from nnuzoo_b200 import _native
    from nnuzoo_b200._native import NzScanDesc

    lib = _native.lib()
    scans = m2net_scan_list(512)
    row_max = max(BATCH * kd * L for kd, L in scans)
    bc_max = max(BATCH * K_DIR * N_STATE * L for _, L in scans)
    pin = lambda n: torch.empty(n, dtype=torch.float32, pin_memory=True)  # noqa: E731
    g = torch.Generator(device=dev).manual_seed(99)
    src_row = torch.randn(row_max, device=dev, generator=g)
    src_bc = torch.randn(bc_max, device=dev, generator=g)
    hu, hd, hgo = pin(row_max), pin(row_max), pin(row_max)
    hB, hC = pin(bc_max), pin(bc_max)
    hu.copy_(src_row)
    hd.copy_(src_row * 0.5)
    hgo.copy_(src_row.flip(0))
    hB.copy_(src_bc)
    hC.copy_(src_bc.flip(0))
    del src_row, src_bc
    ho, hdu, hdd = pin(row_max), pin(row_max), pin(row_max)
    hdB, hdC = pin(bc_max), pin(bc_max)
    kdm = max(kd for kd, _ in scans)
    hA = (-torch.arange(1, N_STATE + 1).float().repeat(kdm, 1)).contiguous().pin_memory()
    hD, hb = torch.ones(kdm).pin_memory(), torch.full((kdm,), -2.0).pin_memory()
    hdA, hdD, hdb = torch.zeros(kdm, N_STATE).pin_memory(), torch.zeros(kdm).pin_memory(), torch.zeros(kdm).pin_memory()
    torch.cuda.synchronize(dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    sp = ctypes.c_void_p(stream.cuda_stream)
    h2d = d2h = 0
    for kd, L in scans:
        b = scan_bytes(BATCH, kd, L)
        h2d += 4 * (3 * b["E"] + 2 * b["S"]) + 4 * kd * (N_STATE + 2)
        d2h += 4 * (3 * b["E"] + 2 * b["S"]) + 4 * kd * (N_STATE + 2)

    def one_step():
        for kd, L in scans:
            d = NzScanDesc()
            d.batch, d.dim, d.dstate, d.ngroups, d.seqlen = BATCH, kd, N_STATE, K_DIR, L
            d.dtype, d.delta_softplus = 0, 1
            d.u, d.delta, d.dout, d.B, d.C = p(hu), p(hd), p(hgo), p(hB), p(hC)
            d.A, d.D, d.delta_bias = p(hA), p(hD), p(hb)
            for s in (d.u_stride, d.delta_stride, d.out_stride, d.dout_stride):
                s[0], s[1] = kd * L, L
            for s in (d.B_stride, d.C_stride):
                s[0], s[1], s[2] = K_DIR * N_STATE * L, N_STATE * L, L
            d.A_stride = N_STATE
            d.out, d.du, d.ddelta, d.dB, d.dC = p(ho), p(hdu), p(hdd), p(hdB), p(hdC)
            d.dA, d.dD, d.ddelta_bias = p(hdA), p(hdD), p(hdb)
            _native.check(lib.nz_scan_fwd_bwd_host(ctypes.byref(d), sp), "nz_scan_fwd_bwd_host")

    for _ in range(warmup):
        one_step()
    torch.cuda.synchronize(dev)
    t0 = time.perf_counter()
    for _ in range(steps):
        one_step()
    torch.cuda.synchronize(dev)
    dt = (time.perf_counter() - t0) / steps
    assert bool(torch.isfinite(ho[:1024]).all())
    return dt, h2d, d2h


def _median_ms(fn, stream, iters, warmup):
    import torch
    for _ in range(warmup):
        fn()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def configs_run(dev, stream, peak):
    """The other BASELINE configs, measured next to the configs[1] headline (parity-test shapes, not bench lines):
    cfg0 = configs[0] (B=2, K=4, D=192, L=4096 fp32, split-view B/C) through the C ABI, fwd and fwd+bwd;
    cfg2 = configs[2]'s 1-D scan (2, 64, 128^3) bf16 + z gate, fwd+bwd through selective_scan_fn;
    cfg3 = configs[3]'s MambaND token counts (75 / 600): microseconds per selective_scan_fn call and launches."""
    import torch

    from nnuzoo_b200 import _native, selective_scan_fn
    from nnuzoo_b200._native import NzScanDesc
    lib = _native.lib()
    sp = ctypes.c_void_p(stream.cuda_stream)
    p = lambda t: ctypes.c_void_p(t.data_ptr())  # noqa: E731
    res = {}
    # ---- cfg0 through the C ABI (kernel time: the Python wrapper would add more than the kernels take) ----
    c = CFG0
    kd, L, N, G, Bn = c["K"] * c["D"], c["L"], c["N"], c["K"], c["batch"]
    u, delta, A, Bv, Cv, D, bias = cfg0_tensors(dev)
    gout = torch.randn(Bn, kd, L, device=dev)
    out, du, dd = torch.empty_like(u), torch.empty_like(u), torch.empty_like(u)
    dB, dC = torch.zeros(Bn, G, N, L, device=dev), torch.zeros(Bn, G, N, L, device=dev)
    dA, dD, db = torch.zeros(kd, N, device=dev), torch.zeros(kd, device=dev), torch.zeros(kd, device=dev)
    x = torch.empty(Bn, kd, (L + _native.NZ_CHUNK - 1) // _native.NZ_CHUNK, N, device=dev)
    d = NzScanDesc()
    d.batch, d.dim, d.dstate, d.ngroups, d.seqlen, d.dtype, d.delta_softplus = Bn, kd, N, G, L, 0, 1
    d.u, d.delta, d.A, d.B, d.C, d.D, d.delta_bias, d.dout = p(u), p(delta), p(A), p(Bv), p(Cv), p(D), p(bias), p(gout)
    for st in (d.u_stride, d.delta_stride, d.out_stride, d.dout_stride):
        st[0], st[1] = kd * L, L
    for k in range(3):
        d.B_stride[k], d.C_stride[k] = Bv.stride(k), Cv.stride(k)
    d.A_stride = N
    d.out, d.x, d.du, d.ddelta = p(out), p(x), p(du), p(dd)
    d.dA, d.dB, d.dC, d.dD, d.ddelta_bias = p(dA), p(dB), p(dC), p(dD), p(db)
    nfine = int(lib.nz_scan_fine_bytes(ctypes.byref(d)))
    xf = torch.empty(max(nfine // 4, 1), device=dev)
    if nfine:
        d.xf = p(xf)
    wsb = max(int(lib.nz_scan_workspace_bytes_bwd(ctypes.byref(d))), int(lib.nz_scan_workspace_bytes_cp(ctypes.byref(d))))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    d.workspace, d.workspace_bytes = p(ws), wsb
    zero = not lib.nz_scan_bwd_overwrites_dbc(ctypes.byref(d))

    def f_fwd():
        _native.check(lib.nz_scan_fwd(ctypes.byref(d), sp), "nz_scan_fwd")

    def f_both():
        _native.check(lib.nz_scan_fwd(ctypes.byref(d), sp), "nz_scan_fwd")
        if zero:
            dB.zero_()
            dC.zero_()
        _native.check(lib.nz_scan_bwd(ctypes.byref(d), sp), "nz_scan_bwd")

    by = scan_bytes(Bn, kd, L)
    t_f, t_fb = _median_ms(f_fwd, stream, 100, 20), _median_ms(f_both, stream, 100, 20)
    res["cfg0"] = {"shape": "B=2, K=4, D=192, N=16, L=4096, fp32, split-view B/C (BASELINE configs[0])",
                   "fwd_us": t_f * 1e3, "fwd_gbps": by["fwd"] / t_f / 1e6, "fwd_frac": by["fwd"] / t_f / 1e6 / peak,
                   "fwd_bwd_us": t_fb * 1e3, "fwd_bwd_gbps": by["total"] / t_fb / 1e6,
                   "fwd_bwd_frac": by["total"] / t_fb / 1e6 / peak, "timing": "median of 100 after 20 warm-ups, CUDA events, C ABI",
                   "note": "warm in L2 (214 MB of operands + results vs 126 MB L2); launch-latency bound"}
    del u, delta, Bv, Cv, gout, out, du, dd, dB, dC, x, xf, ws
    # ---- cfg2: the 1-D nets' scan ----
    torch.manual_seed(3)
    Bn, Dm, L = 2, 64, 128 ** 3
    mk = lambda *sh: torch.randn(*sh, device=dev).to(torch.bfloat16)  # noqa: E731
    leaves = [mk(Bn, Dm, L).requires_grad_(True), (0.5 * torch.randn(Bn, Dm, L, device=dev)).to(torch.bfloat16).requires_grad_(True),
              mk(Bn, 1, 16, L).requires_grad_(True), mk(Bn, 1, 16, L).requires_grad_(True), mk(Bn, Dm, L).requires_grad_(True)]
    A = (-torch.arange(1, 17, device=dev).float().repeat(Dm, 1)).requires_grad_(True)
    Dp, bias = torch.ones(Dm, device=dev, requires_grad=True), torch.full((Dm,), -2.0, device=dev, requires_grad=True)
    g2 = mk(Bn, Dm, L)

    def f2_fwd():
        with torch.no_grad():
            selective_scan_fn(leaves[0], leaves[1], A, leaves[2], leaves[3], Dp, leaves[4], bias, True)

    def f2():
        o = selective_scan_fn(leaves[0], leaves[1], A, leaves[2], leaves[3], Dp, leaves[4], bias, True)
        o.backward(g2)
        for t in leaves + [A, Dp, bias]:
            t.grad = None

    E, S = Bn * Dm * L, Bn * 16 * L
    t2f, t2 = _median_ms(f2_fwd, stream, 5, 2), _median_ms(f2, stream, 5, 2)
    res["cfg2"] = {"shape": "u (2, 64, 2097152) bf16, z gate, one B/C group (BASELINE configs[2]'s 1-D scan)",
                   "fwd_ms": t2f, "fwd_gbps": 2 * (4 * E + 2 * S) / t2f / 1e6,
                   "fwd_bwd_ms": t2, "fwd_bwd_gbps": 2 * (11 * E + 6 * S) / t2 / 1e6,
                   "fwd_bwd_frac": 2 * (11 * E + 6 * S) / t2 / 1e6 / peak,
                   "timing": "median of 5 after 2 warm-ups through selective_scan_fn (autograd included)"}
    del leaves, g2
    torch.cuda.empty_cache()
    # ---- cfg3: MambaND token counts ----
    rows = []
    for Dm, L in ((192, 600), (384, 600), (768, 75)):
        lv = [torch.randn(2, Dm, L, device=dev, requires_grad=True), (0.5 * torch.randn(2, Dm, L, device=dev)).requires_grad_(True),
              torch.randn(2, 1, 16, L, device=dev, requires_grad=True), torch.randn(2, 1, 16, L, device=dev, requires_grad=True),
              torch.randn(2, Dm, L, device=dev, requires_grad=True)]
        A = (-torch.arange(1, 17, device=dev).float().repeat(Dm, 1)).requires_grad_(True)
        Dp, bias = torch.ones(Dm, device=dev, requires_grad=True), torch.full((Dm,), -2.0, device=dev, requires_grad=True)
        g3 = torch.randn(2, Dm, L, device=dev)

        def f3f():
            with torch.no_grad():
                selective_scan_fn(lv[0], lv[1], A, lv[2], lv[3], Dp, lv[4], bias, True)

        def f3():
            o = selective_scan_fn(lv[0], lv[1], A, lv[2], lv[3], Dp, lv[4], bias, True)
            o.backward(g3)

        n0 = _native.launch_count()
        f3()
        per_call = _native.launch_count() - n0
        row = {"shape": [2, Dm, L], "fwd_us": _median_ms(f3f, stream, 50, 10) * 1e3,
               "fwd_bwd_us": _median_ms(f3, stream, 50, 10) * 1e3, "our_launches_fwd_bwd": per_call}
        try:  # the same calls replayed from a CUDA graph: device time without the Python / allocator / launch overhead
            for t in lv + [A, Dp, bias]:
                t.grad = None
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    f3f()
                    o = selective_scan_fn(lv[0], lv[1], A, lv[2], lv[3], Dp, lv[4], bias, True)
                    torch.autograd.grad(o, lv + [A, Dp, bias], g3)
            torch.cuda.current_stream(dev).wait_stream(side)
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf):
                f3f()
            with torch.cuda.graph(gb):
                o = selective_scan_fn(lv[0], lv[1], A, lv[2], lv[3], Dp, lv[4], bias, True)
                keep = torch.autograd.grad(o, lv + [A, Dp, bias], g3)
            cur = torch.cuda.current_stream(dev)
            row["graph_fwd_us"] = _median_ms(gf.replay, cur, 100, 20) * 1e3
            row["graph_fwd_bwd_us"] = _median_ms(gb.replay, cur, 100, 20) * 1e3
            del keep, gf, gb
        except Exception as e:  # noqa: BLE001
            row["graph_error"] = f"{type(e).__name__}: {e}"[:200]
        rows.append(row)
    res["cfg3"] = {"what": "BASELINE configs[3] (MambaND2Net) token counts: microseconds per selective_scan_fn call "
                           "(Python wrapper + allocations + launches), fp32, z gate", "calls": rows}
    del lv, g3
    torch.cuda.empty_cache()
    # ---- cfg2 as a config, not a shape: one ResMambaBlock (lm2net.py:107-176) of the stage-1 LightM-UNet level ----
    try:
        res["cfg2"]["block"] = _resmamba_block_run(dev, stream)
    except Exception as e:  # noqa: BLE001  (a side measurement must not take the headline line down)
        res["cfg2"]["block"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    # ---- cfg3 as a config: the MambaNDCore block stack (mamba_nd2net.py:725-1001) ----
    try:
        res["cfg3"]["core"] = _mamba_nd_core_run(dev, stream)
    except Exception as e:  # noqa: BLE001
        res["cfg3"]["core"] = {"error": f"{type(e).__name__}: {e}"[:300]}
    torch.cuda.empty_cache()
    return res


def _resmamba_block_run(dev, stream):
    """BASELINE configs[2]: ResMambaBlock(spatial_dims=3, in_channels=32, GroupNorm(8)) on a (2, 32, 128, 128, 128)
    feature map under bf16 autocast, forward + backward, for the three axis orders LM2Net cycles through
    (lm2net.py:286-315).  Each block runs two MambaLayers = two (2, 64, 2 097 152) scans with z gate."""
    import torch

    from nnuzoo_b200 import _native
    from nnuzoo_b200.mamba_nd import ResMambaBlock
    torch.manual_seed(11)
    x = torch.randn(2, 32, 128, 128, 128, device=dev, requires_grad=True)
    out = {"shape": "x (2, 32, 128, 128, 128), bf16 autocast, ResMambaBlock fwd+bwd (2 MambaLayers, d_inner 64, "
                    "L = 2 097 152 each)", "orders": {}}
    for order in ("d h w", "d w h", "w h d"):
        blk = ResMambaBlock(3, 32, norm=("GROUP", {"num_groups": 8}), order=order).to(dev)

        def f():
            with torch.autocast("cuda", dtype=torch.bfloat16):
                y = blk(x)
            y.float().square().mean().backward()
            x.grad = None
            blk.zero_grad(set_to_none=True)

        n0 = _native.launch_count()
        f()
        launches = _native.launch_count() - n0
        out["orders"][order] = {"fwd_bwd_ms": _median_ms(f, stream, 3, 1), "our_launches": launches}
        del blk
        torch.cuda.empty_cache()
    out["timing"] = "median of 3 after 2 warm-ups, CUDA events, whole block (GSC convolutions / norms are library kernels)"
    return out


def _mamba_nd_core_run(dev, stream):
    """BASELINE configs[3]: the MambaNDCore block stack at MambaND2Net's token grid -- 7 Blocks, d_model 96, tokens
    (6, 10, 10), batch 2, three orders x reversed odd layers (mamba_nd2net.py:960-1001) -- eager and as a replayed CUDA
    graph (the path is launch-bound: L = 600)."""
    import torch

    from nnuzoo_b200 import _native
    from nnuzoo_b200.mamba_nd import MambaNDCore
    torch.manual_seed(12)
    nl = 7
    core = MambaNDCore(spatial_dims=3, img_size=(12, 20, 20), patch_size=(2, 2, 2), in_channels=1, embed_dims=96,
                       num_layers=nl, fused_add_norm=False, final_norm=False).to(dev)
    shape = (6, 10, 10)
    x = torch.randn(2, 600, 96, device=dev, requires_grad=True)
    gy = torch.randn(2, 600, 96, device=dev)

    def fwd():
        with torch.no_grad():
            core.forward_tokens(x, shape)

    def both():
        y, _ = core.forward_tokens(x, shape)
        y.backward(gy)
        x.grad = None
        core.zero_grad(set_to_none=True)

    n0 = _native.launch_count()
    both()
    launches = _native.launch_count() - n0
    res = {"shape": "tokens (2, 600, 96) = grid (6, 10, 10), 7 Blocks, d_inner 192, fp32",
           "fwd_us_per_block": _median_ms(fwd, stream, 30, 5) * 1e3 / nl,
           "fwd_bwd_us_per_block": _median_ms(both, stream, 30, 5) * 1e3 / nl,
           "our_launches_per_block_fwd_bwd": launches / nl}
    # replayed CUDA graph of forward + backward (static input / gradient buffers)
    try:
        params = [q for q in core.parameters()]
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(3):
                y, _ = core.forward_tokens(x, shape)
                torch.autograd.grad(y, [x] + params, gy, allow_unused=True)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            y, _ = core.forward_tokens(x, shape)
            grads = torch.autograd.grad(y, [x] + params, gy, allow_unused=True)
        _ = grads
        t = _median_ms(graph.replay, torch.cuda.current_stream(dev), 50, 10)
        res["graph_fwd_bwd_us_per_block"] = t * 1e3 / nl
    except Exception as e:  # noqa: BLE001
        res["graph_error"] = f"{type(e).__name__}: {e}"[:300]
    return res


def train_run(dev, world, steps, warmup, per_gpu_batch, sync_bn=True, scaling="weak"):
    """The second half of BASELINE.json's metric: SS2D2Net (M2Net) training patches/s -- full optimisation steps
    (H2D of the pinned batch, bf16-autocast forward, Dice+CE deep-supervision loss, backward with the DDP gradient
    all-reduce over NCCL, clip, AdamW, loss read back) through nnuzoo_b200.train.Trainer.  Weak scaling: every rank
    trains ``per_gpu_batch`` patches of 1x512x512, like the scan part of this line."""
    import torch

    from nnuzoo_b200 import _native
    from nnuzoo_b200.m2net import get_m2net
    from nnuzoo_b200.train import Trainer, synthetic_batch

    torch.manual_seed(0)
    trainer = Trainer(get_m2net(1, 4, True).train(), dev, sync_bn=sync_bn)   # SyncBatchNorm as nnUNetTrainer.py:279-280
    data, targets = synthetic_batch(per_gpu_batch, 1, 4, seed=17 + dev.index)
    h2d = data.numel() * data.element_size() + sum(t.numel() * t.element_size() for t in targets)
    for _ in range(warmup):
        trainer.train_step(data, targets)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize(dev)
    n0 = _native.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    losses = torch.zeros(steps, device=dev)
    for i in range(steps):
        losses[i] = trainer.train_step(data, targets)             # the step's result stays on the device ...
    loss = float(losses[-1].item())                                # ... and is read back once (no host sync per step)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = max_over_ranks(e0.elapsed_time(e1) / steps, dev, world)
    if world > 1:
        gb = torch.tensor([float(per_gpu_batch)], device=dev)
        torch.distributed.all_reduce(gb)
        global_batch = int(gb.item())
    else:
        global_batch = per_gpu_batch
    return {"patches_per_s": global_batch / (ms * 1e-3), "unit": "patches/s", "ms_per_step": ms,
            "per_gpu_batch": per_gpu_batch, "global_batch": global_batch, "scaling": scaling, "sync_bn": bool(sync_bn and world > 1),
            "steps": steps, "warmup": warmup, "autocast": "bf16", "loss": loss,
            "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
            "our_kernel_launches_per_step": (_native.launch_count() - n0) // steps,
            "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9,
            "what": "M2Net(1->4 classes, deep supervision) full training step via nnuzoo_b200.train.Trainer; "
                    "DDP (NCCL all-reduce) when n_gpus > 1"}


def infer_run(dev, world, slices, tile_batch):
    """BASELINE configs[4]: sliding-window inference of SS2D2Net over a synthetic (1, slices, 512, 512) volume, 2-D
    tiles of 512x512 (step 0.5), mirroring over both axes (4 forwards per tile, stacked), gaussian-weighted fp16
    accumulators; the 200 disjoint slices are dealt as contiguous balanced runs per rank and merged with one all-gather
    of the finished slabs (strong scaling: one volume whatever N).  Timed: volume already on the device -> merged logits on the device."""
    import torch

    from nnuzoo_b200.m2net import get_m2net
    from nnuzoo_b200.predict import SlidingWindowPredictor

    torch.manual_seed(0)
    net = get_m2net(1, 4, False)
    pred = SlidingWindowPredictor(net, (512, 512), 4, dev, tile_batch=tile_batch, autocast_dtype=torch.bfloat16)
    g = torch.Generator().manual_seed(5)
    vol = torch.randn(1, slices, 512, 512, generator=g).to(dev)
    pred.predict_logits(vol[:, :2 * tile_batch * world])   # warm-up (allocator, cuDNN algorithm choice)
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
        torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = pred.predict_logits(vol)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    return {"volumes_per_s": 1e3 / ms, "slices_per_s": slices * 1e3 / ms, "seconds_per_volume": ms / 1e3,
            "volume": [1, slices, 512, 512], "tile_batch": tile_batch, "forwards_per_rank": pred.forwards,
            "mirroring": "axes (0, 1), 4 passes per tile stacked into one forward", "autocast": "bf16 (random-init weights overflow the reference's fp16 default)",
            "scaling": "strong", "finite": bool(torch.isfinite(out).all()),
            "merge": ("disjoint slices: contiguous balanced runs per rank, divided locally, ONE all_gather of the finished fp16 "
                      "slabs over NCCL (nnuzoo_b200/predict.py)") if world > 1 else "single rank",
            "accumulate": "nz_sw_accumulate (mirror average + gaussian multiply-accumulate, one launch per tile batch)"}


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: this path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        import datetime
        # a collective that cannot complete (a rank gone, mismatched sizes) aborts after 5 minutes instead of NCCL's 10
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
    from nnuzoo_b200 import _native

    wl = Workload(dev)
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    min_warm = 0 if os.environ.get("NZ_BENCH_PROFILE") else 3  # profiling runs may skip the warm-up
    for _ in range(max(min_warm, args.warmup)):
        wl.step(stream)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    n0 = _native.launch_count()
    bwd_events = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        wl.step(stream, bwd_events)
    e1.record(stream)
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = _native.launch_count() - n0
    ms = max_over_ranks(e0.elapsed_time(e1), dev, world)
    ms_per_step = ms / args.steps
    value = world * wl.total_bytes / (ms_per_step * 1e-3) / 1e9
    bwd_ms = sum(a.elapsed_time(b) for a, b in bwd_events)
    bwd_gbps = wl.bwd_bytes * args.steps / (bwd_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    e2e = None
    cpu = None
    if not args.no_e2e:
        try:
            dt, h2d, d2h = e2e_run(dev, stream, args.e2e_steps, 1)
            barrier()
            dt = max_over_ranks(dt, dev, world)
            e2e = {"value": world * wl.total_bytes / dt / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "steps": args.e2e_steps, "ms_per_step": dt * 1e3,
                   "api": "nz_scan_fwd_bwd_host (include/nnuzoo_b200.h), pinned host buffers"}
        except Exception as ex:  # report, never fake
            e2e = {"value": None, "unit": "GB/s", "error": repr(ex)[:300]}
    configs = None
    if rank == 0 and not args.no_configs:
        try:
            configs = configs_run(dev, stream, peak)
        except Exception as ex:  # report, never fake
            configs = {"error": repr(ex)[:300]}
    train = None
    deferred_strong = None
    if not args.no_train:
        del wl.u, wl.delta, wl.dout, wl.Bm, wl.Cm, wl.out, wl.du, wl.dd, wl.dB, wl.dC, wl.x, wl.xf, wl.ws
        torch.cuda.empty_cache()
        try:
            train = train_run(dev, world, args.train_steps, 2, args.train_batch)
            if world > 1:  # the reference's own split of a global batch of 12 (nnUNetTrainer.py:420-429): strong scaling
                from nnuzoo_b200.train import split_global_batch
                sizes = split_global_batch(BATCH, world)
                if len(set(sizes)) == 1:
                    torch.cuda.empty_cache()
                    train["strong"] = train_run(dev, world, args.train_steps, 2, sizes[rank],
                                                scaling="strong (global batch 12 split as the reference does)")
                else:
                    # an uneven split (8 ranks: 2,2,2,2,1,1,1,1) hung this leg once (per-sample batch-dice all-gather,
                    # fixed in nnuzoo_b200/train.py); it is therefore measured AFTER the line below is out, best effort
                    # and under a hard exit timer, so it can never take the measured line down with it
                    deferred_strong = sizes
                    train["strong"] = {"deferred": f"uneven split {sizes}: measured after this line is printed and "
                                                   "reported on stderr as 'STRONG_SCALING {json}'"}
        except Exception as ex:  # report, never fake
            train = {"patches_per_s": None, "error": repr(ex)[:300]}
    infer = None
    if not args.no_infer:
        torch.cuda.empty_cache()
        try:
            infer = infer_run(dev, world, args.infer_slices, args.infer_tile_batch)
        except Exception as ex:  # report, never fake
            infer = {"volumes_per_s": None, "error": repr(ex)[:300]}
    if rank == 0 and world == 1 and not args.no_cpu:
        r = cpu_sample_run(1, 1)
        cpu = {"value": r["gbps"], "unit": "GB/s", "cores": r["cores"], "kind": "port", "sample": r["sample"],
               "seconds": r["seconds"]}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": CONFIG,
            "patches_per_s_scan_only": world * BATCH / (ms_per_step * 1e-3),
            "frac_of_hbm_peak": value / world / peak,
            "roofline": {"bound": "hbm", "achieved": bwd_gbps, "peak": peak, "unit": "GB/s", "frac": bwd_gbps / peak,
                         "traffic": traffic_per_launch()[0], "traffic_git": traffic_per_launch()[1], "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": wl.bwd_bytes / len(wl.scans),
                         "kernel": "nz_scan_bwd: from 12 M elements the row-per-lane backward (nz::scan_rl_agg_kernel + "
                                   "scan_rl_combine_kernel + nz::scan_bwd_rl2_kernel, csrc/scan_rl_kernels.cuh), below that "
                                   "nz::scan_bwd_kernel<float, 8, 16, 8, TMA>; all 80 calls per step; algorithmic bytes "
                                   "4*(5E+4S) per call, achieved = sum of bytes / sum of CUDA-event durations of the calls",
                         "share_of_step": bwd_ms / (ms_per_step * args.steps),
                         "mufu_floor_frac": 0.54,
                         "mufu_floor_note": "16 ex2 per element and pass at 16 lanes/clk/SM: the row-per-lane backward "
                                            "evaluates them twice (aggregate pass + main pass, 2.3 clk/elt) against an HBM "
                                            "floor of 1.24 clk/elt at the stage-1 shape, so MUFU caps it at 0.54 of the "
                                            "measured HBM peak (the single-pass warp-scan kernel at 0.90)"},
            "configs": configs, "train": train, "infer": infer, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(line), flush=True)
    if deferred_strong is not None:
        import threading
        guard = threading.Timer(240.0, lambda: os._exit(0))   # a collective that cannot complete ends the run cleanly
        guard.daemon = True
        guard.start()
        try:
            torch.cuda.empty_cache()
            strong = train_run(dev, world, args.train_steps, 2, deferred_strong[rank],
                               scaling="strong (global batch 12 split as the reference does)")
            if rank == 0:
                print("STRONG_SCALING " + json.dumps(strong), file=sys.stderr, flush=True)
        except Exception as ex:  # noqa: BLE001
            if rank == 0:
                print("STRONG_SCALING failed: " + repr(ex)[:300], file=sys.stderr, flush=True)
        guard.cancel()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0] / [2] / [3] side measurements")
    ap.add_argument("--infer-slices", type=int, default=200)
    ap.add_argument("--infer-tile-batch", type=int, default=5)
    ap.add_argument("--train-steps", type=int, default=10)
    ap.add_argument("--train-batch", type=int, default=BATCH, help="per-GPU batch of the M2Net training leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
