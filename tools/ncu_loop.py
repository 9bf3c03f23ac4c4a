"""Per-trip instruction mix of the hot loop from an `ncu --page source --csv` dump.

    python tools/ncu_loop.py CSV TRIPS [dump]
TRIPS = executions of the loop body per warp summed over the launch (e.g. warp-tiles * states).
Instructions executed >= 0.2*TRIPS count as loop body, the rest as per-tile code."""
import collections
import csv
import sys


def main(path, trips, dump=False):
    rows = list(csv.reader(open(path)))
    hdr, start = (rows[1], 2) if rows[0][0].startswith("Kernel") else (rows[0], 1)
    ix, iex, isamp, ia = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Address")
    seen = set()
    ops, samp = collections.Counter(), collections.Counter()
    out_ops, out_samp = collections.Counter(), 0
    lines = []
    for r in rows[start:]:
        try:
            c = int(r[iex])
        except (ValueError, IndexError):
            continue
        if r[ia] in seen:
            continue
        seen.add(r[ia])
        s = r[ix].split()
        op = s[1] if s[0].startswith("@") else s[0]
        op = ".".join(op.split(".")[:2])
        sm = int(r[isamp]) if r[isamp].isdigit() else 0
        if c >= 0.2 * trips:
            ops[op] += c / trips
            samp[op] += sm
            lines.append("%s %5.2f %5d  %s" % (r[ia][-5:], c / trips, sm, r[ix]))
        elif c > 0:
            out_ops[op] += c / trips
            out_samp += sm
    print("loop: %.1f instr per trip; outside the loop: %.1f per trip (%d stall samples outside, %d inside)"
          % (sum(ops.values()), sum(out_ops.values()), out_samp, sum(samp.values())))
    for k, v in ops.most_common(40):
        print("  %-18s %6.1f  samples %6d" % (k, v, samp[k]))
    print("outside:", "  ".join("%s:%.1f" % kv for kv in out_ops.most_common(16)))
    if dump:
        print("\n".join(lines))


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]), len(sys.argv) > 3)
