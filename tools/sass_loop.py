"""Static op histogram of the innermost-large loop of a kernel in an object file.

    python tools/sass_loop.py OBJ KERNEL_SUBSTR [lo hi]   (addresses in hex to override loop detection)
"""
import collections
import re
import subprocess
import sys


def main():
    obj, key = sys.argv[1], sys.argv[2]
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    funcs = re.split(r"\n\s*Function : ", out)
    body = next(f for f in funcs if key in f.split("\n")[0])
    ins = []
    for l in body.split("\n"):
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", l)
        if m:
            ins.append((int(m.group(1), 16), m.group(2).strip()))
    if len(sys.argv) > 4:
        lo, hi = int(sys.argv[3], 16), int(sys.argv[4], 16)
    else:
        # largest backward branch that is not the outermost one
        loops = []
        for a, s in ins:
            m = re.search(r"BRA(?:\.U)?(?:\s+!?U?P\d,)?\s+(0x[0-9a-f]+)", s)
            if m and int(m.group(1), 16) < a:
                loops.append((a - int(m.group(1), 16), int(m.group(1), 16), a))
        loops.sort(reverse=True)
        print("loops (bytes, lo, hi):", [(b, hex(l), hex(h)) for b, l, h in loops[:4]])
        _, lo, hi = loops[1] if len(loops) > 1 else loops[0]
    sel = [(a, s) for a, s in ins if lo <= a <= hi]
    c = collections.Counter()
    for a, s in sel:
        t = s.split()
        op = t[1] if t[0].startswith("@") else t[0]
        parts = op.split(".")
        name = parts[0]
        if name in ("SHFL", "LDS", "STS", "MUFU", "BAR", "REDG", "ATOMG", "STG", "LDG") and len(parts) > 1:
            name += "." + parts[1]
        c[name] += 1
    print(f"{key}: loop {hex(lo)}..{hex(hi)}: {len(sel)} instructions")
    print("  " + "  ".join(f"{k}:{v}" for k, v in c.most_common(40)))


if __name__ == "__main__":
    main()
