"""Build kernel variants (compile-time switches of scan_kernels.cuh) as separate libraries under
gpurun_out/variants/ for A/B timing on the GPU box:  python tools/tune_build.py name:-DX=1,-DY=2 ..."""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "nnuzoo_b200", "csrc")
OUT = os.path.join(ROOT, "tune_variants")
BASE = ["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
        "-Xcompiler", "-fPIC", "-ccbin", "/usr/bin/g++", "-I", os.path.join(ROOT, "include")]
sys.path.insert(0, ROOT)
from nnuzoo_b200.build import SOURCES as SRCS  # noqa: E402  (every translation unit of the library)


def build(spec):
    name, _, defs = spec.partition(":")
    flags = [d for d in defs.split(",") if d] + ["-DNZ_F32_ONLY"]
    d = os.path.join(OUT, name)
    os.makedirs(d, exist_ok=True)

    def cc(src):
        obj = os.path.join(d, src.replace(".cu", ".o"))
        r = subprocess.run(BASE + flags + ["-Xptxas", "-v", "-c", os.path.join(CSRC, src), "-o", obj],
                           capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stderr)
            raise SystemExit(f"{name}: nvcc failed on {src}")
        if src == "scan_inst_f32.cu":
            regs = [l for l in r.stderr.splitlines() if "Compiling entry" in l or "registers" in l or "spill" in l]
            with open(os.path.join(d, "ptxas_f32.txt"), "w") as f:
                f.write("\n".join(regs))
        return obj

    with ThreadPoolExecutor(5) as ex:
        objs = list(ex.map(cc, SRCS))
    lib = os.path.join(d, "libnnuzoo_b200.so")
    subprocess.check_call(BASE[:1] + ["-shared", "-o", lib, *objs, "-ccbin", "/usr/bin/g++"])
    return lib


if __name__ == "__main__":
    for spec in sys.argv[1:]:
        print(build(spec))
