"""Summary of every launch in an .ncu-rep (--set full): duration, DRAM bytes, pipes, occupancy, achieved residency and
the stall reasons (warps per issue-active cycle).

    python tools/ncu_stalls.py REPORT.ncu-rep [kernel-substring]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__occupancy_limit_shared_mem", "CTA/SM limit (smem)"),
    ("launch__occupancy_limit_registers", "CTA/SM limit (regs)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("sm__inst_issued.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU (MUFU) pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / scheduler / cycle"),
    ("lts__t_sectors_op_red.sum", "L2 RED sectors"),
    ("lts__t_sectors_op_write.sum", "L2 write sectors"),
    ("lts__t_sectors_op_read.sum", "L2 read sectors"),
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    flt = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr, units = rows[hi], rows[hi + 1]
    for r in rows[hi + 2:]:
        if len(r) < len(hdr):
            continue
        name = r[hdr.index("Kernel Name")]
        if flt not in name:
            continue
        print("=" * 100)
        print(name[:150])
        for k, label in KEYS:
            if k in hdr:
                print(f"  {label:38s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")
        st = []
        for i, h in enumerate(hdr):
            if h.startswith(STALLS) and h.endswith("_per_warp_active.pct") is False and "not_issued" not in h:
                try:
                    st.append((float(r[i].replace(",", "")), h[len(STALLS):].replace("_per_warp_active.pct", "").replace(".ratio", "")))
                except ValueError:
                    pass
        st.sort(reverse=True)
        print("  stalls: " + ", ".join(f"{n} {v:.2f}" for v, n in st[:9]))


if __name__ == "__main__":
    main()
