"""Summarise an `ncu --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum`
launch list of `bench.py --steps 1`:

  * per-kernel share of the summed launch time (compare with bench.py's CUDA-event share; ncu's
    absolute times are cold-cache and serialised),
  * average DRAM traffic per scan_bwd launch next to the average algorithmic bytes per launch
    -> profiles/r01_bwd_traffic.json, which bench.py reports as roofline.traffic.

    python tools/ncu_traffic.py gpurun_out/launches_r01.csv [profiles/r01_bwd_traffic.json]
"""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main(path, out=None):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 10]
    hdr = rows[0]
    ik, im, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iid = hdr.index("ID")
    per = collections.OrderedDict()
    for r in rows[1:]:
        d = per.setdefault(r[iid], {"kernel": r[ik]})
        v = float(r[iv].replace(",", ""))
        unit = r[iu]
        if r[im].startswith("gpu__time_duration"):
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)  # -> us
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)  # -> bytes
        d[r[im]] = v
    tot = sum(d.get("gpu__time_duration.sum", 0.0) for d in per.values())
    by = collections.Counter()
    cnt = collections.Counter()
    for d in per.values():
        name = d["kernel"].split("<")[0].split("(")[0]
        by[name] += d.get("gpu__time_duration.sum", 0.0)
        cnt[name] += 1
    print(f"{len(per)} launches, {tot / 1e3:.2f} ms summed (serialised, cold cache)")
    for k, v in by.most_common():
        print(f"  {k:40s} launches {cnt[k]:4d}  time {v / 1e3:9.3f} ms  share {v / tot:6.1%}")
    # every launch an nz_scan_bwd call makes: the warp-scan kernel, or the row-per-lane trio (backward aggregate pass =
    # scan_rl_agg_kernel<T, z, false>, its combine, main pass)
    order = list(per.values())
    bwd = []
    for i, d in enumerate(order):
        k = d["kernel"]
        kk = k.replace(" ", "")
        # template arguments of the aggregate kernel: <T, kHasZ, kFwd, kRevCap>; the backward's passes have kFwd = 0
        is_bwd_agg = False
        if "scan_rl_agg_kernel<" in kk:
            targs = kk.split("scan_rl_agg_kernel<", 1)[1].split(">", 1)[0].split(",")
            is_bwd_agg = len(targs) >= 3 and targs[2].replace("(bool)", "") in ("0", "false")
        is_bwd_comb = "combine" in k and i > 0 and order[i - 1] in bwd and "agg" in order[i - 1]["kernel"]
        if "scan_bwd" in k or is_bwd_agg or is_bwd_comb:
            bwd.append(d)
    if bwd and "dram__bytes_read.sum" in bwd[0]:
        scans = bench.m2net_scan_list(512)
        calls = len(scans)
        traffic = sum(d["dram__bytes_read.sum"] + d["dram__bytes_write.sum"] for d in bwd) / calls
        algo = sum(bench.scan_bytes(bench.BATCH, kd, L)["bwd"] for kd, L in scans) / calls
        t_bwd = sum(d.get("gpu__time_duration.sum", 0.0) for d in bwd)
        print(f"nz_scan_bwd: {len(bwd)} launches in {calls} calls, {t_bwd / 1e3:.2f} ms ({t_bwd / tot:.1%} of the step), DRAM "
              f"traffic {traffic / 1e6:.1f} MB per call vs algorithmic {algo / 1e6:.1f} MB per call (x{traffic / algo:.3f})")
        if out:
            import subprocess
            try:
                rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
            except Exception:
                rev = None
            json.dump({"dram_bytes_per_launch": traffic, "algorithmic_bytes_per_launch": algo, "launches": len(bwd),
                       "calls": calls, "ratio": traffic / algo, "share_of_step_ncu": t_bwd / tot, "git": rev,
                       "source": os.path.basename(path), "unit": "per nz_scan_bwd call (1 or 3 launches)",
                       "how": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum "
                              "--clock-control none, bench.py --steps 1 (NZ_BENCH_PROFILE=1)"}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
