"""Summarise an `ncu --page source --csv` dump: hot-loop instruction mix and stall reasons."""
import collections
import csv
import sys


def main(path, frac=0.4, show=0):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    ix, ie, isamp = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    stall_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    data = []
    for r in rows[2:]:
        if len(r) <= ie:
            continue
        try:
            c = int(r[ie])
        except ValueError:
            c = 0
        data.append((c, r[ix], r))
    tot = sum(c for c, _, _ in data)
    mx = max(c for c, _, _ in data)
    hot = [(c, s, r) for c, s, r in data if c >= mx * frac]
    print(f"total warp-instr {tot}; hottest count {mx}; {len(hot)} instrs >= {frac}*max "
          f"(cover {sum(c for c, _, _ in hot) / tot:.1%} of all)")
    ops = collections.Counter()
    for c, s, _ in hot:
        toks = s.split()
        op = toks[1] if toks[0].startswith("@") else toks[0]
        ops[".".join(op.split(".")[:2])] += c
    print("hot-region op mix (executions / hottest count = per loop trip):")
    for k, v in ops.most_common(40):
        print(f"  {k:22s} {v / mx:8.1f}")
    st = collections.Counter()
    nsamp = 0
    for c, s, r in data:
        for i, h in stall_cols:
            try:
                st[h] += int(r[i])
            except (ValueError, IndexError):
                pass
        try:
            nsamp += int(r[isamp])
        except ValueError:
            pass
    print("stall samples:", nsamp)
    for k, v in st.most_common(12):
        print(f"  {k:26s} {v:8d} {v / max(1, sum(st.values())):6.1%}")
    if show:
        top = sorted(data, key=lambda t: -(int(t[2][isamp]) if t[2][isamp].isdigit() else 0))[:show]
        print("top sampled instructions:")
        for c, s, r in top:
            print(f"  samples {r[isamp]:>6s} exec {c:>10d}  {s[:110]}")


if __name__ == "__main__":
    main(sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 0.4, int(sys.argv[3]) if len(sys.argv) > 3 else 0)
