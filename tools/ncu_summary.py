"""One-screen summary of an .ncu-rep (--set full) for the scan kernels: duration, DRAM bytes and
throughput, pipe utilisation, occupancy, top stall reasons and the hot loop's instruction mix.

    python tools/ncu_summary.py REPORT.ncu-rep TRIPS_FWD TRIPS_BWD > profiles/r01_xxx.txt
TRIPS = warp-level executions of the state loop body in one launch (warp-tiles x states)."""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active % (occupancy)"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "shared-memory wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "shared-memory bank conflicts"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main(rep, trips_f, trips_b):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print("=" * 100)
        print(name[:140])
        for k, label in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {label:42s} {r[i]:>16s} {units[i]}")
        st = sorted(((float(r[i]), h[len(STALLS):-len('_per_issue_active.ratio')]) for i, h in enumerate(hdr)
                     if h.startswith(STALLS) and h.endswith("_per_issue_active.ratio")), reverse=True)
        print("  stalled warps per issue-active cycle:", ", ".join(f"{n} {v:.2f}" for v, n in st[:7]))
    for pat, trips in (("scan_fwd", trips_f), ("scan_bwd", trips_b)):
        src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + pat],
                             capture_output=True, text=True).stdout
        tmp = f"/tmp/_src_{pat}.csv"
        open(tmp, "w").write(src)
        print("=" * 100)
        print(f"{pat}: hot loop (state loop) instruction mix per trip, {trips:.0f} trips")
        out = subprocess.run([sys.executable, __file__.replace("ncu_summary", "ncu_loop"), tmp, str(trips)],
                             capture_output=True, text=True).stdout
        print("\n".join(out.splitlines()[:34]))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), float(sys.argv[3]))
