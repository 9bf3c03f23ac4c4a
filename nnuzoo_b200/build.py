"""Build the native library in-tree: nnuzoo_b200/lib/libnnuzoo_b200.so (sm_100a only).

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  ``python -m nnuzoo_b200.build [--force] [-v]``.  Objects are rebuilt only when
their source, a header they include, the C ABI header or the flags changed.
"""
from __future__ import annotations

import hashlib
import os
import re
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIBDIR = os.path.join(_HERE, "lib")
OBJDIR = os.path.join(LIBDIR, "obj")
LIB = os.path.join(LIBDIR, "libnnuzoo_b200.so")
INCLUDE = os.path.join(os.path.dirname(_HERE), "include")

SOURCES = ["capi.cu", "cross_kernels.cu", "conv1d_kernels.cu", "proj_kernels.cu", "norm_kernels.cu", "dwconv_kernels.cu",
           "epilogue_kernels.cu", "sw_kernels.cu", "scan_inst_f32.cu", "scan_inst_bf16.cu", "scan_inst_f16.cu", "scan_rl_inst_f32.cu",
           "scan_rl_inst_bf16.cu", "scan_rl_inst_f16.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xptxas=-warn-spills"]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _deps(path: str, seen=None) -> list:
    """The file and every local header it includes (transitively), sorted."""
    seen = set() if seen is None else seen
    if path in seen or not os.path.exists(path):
        return []
    seen.add(path)
    with open(path) as fh:
        text = fh.read()
    for inc in re.findall(r'#include\s+"([^"]+)"', text):
        _deps(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)
    return sorted(seen)


def _src_stamp(src: str) -> str:
    h = hashlib.sha256()
    for f in _deps(os.path.join(CSRC, src)):
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stamp() -> str:
    h = hashlib.sha256()
    for s in SOURCES:
        h.update(_src_stamp(s).encode())
    return h.hexdigest()


def is_current() -> bool:
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    return os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read().strip() == _stamp()


def build_native(force: bool = False, verbose: bool = False) -> str:
    """Compile every CUDA source for sm_100a and link the shared library.  Returns its path."""
    if not force and is_current():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    nvcc = _nvcc()
    env = dict(os.environ)
    # the image exports CC/CXX pointing at a gcc without libgomp specs; nvcc only needs a host g++
    ccbin = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else None

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJDIR, src.replace(".cu", ".o"))
        stamp_file, stamp = obj + ".stamp", _src_stamp(src)
        if not force and os.path.exists(obj) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", os.path.join(CSRC, src), "-o", obj]
        if ccbin:
            cmd += ["-ccbin", ccbin]
        r = subprocess.run(cmd, capture_output=True, text=True, env=env)
        if verbose or r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
        with open(stamp_file, "w") as f:
            f.write(stamp)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
    if ccbin:
        cmd += ["-ccbin", ccbin]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if r.returncode:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("link failed")
    with open(os.path.join(LIBDIR, "build.stamp"), "w") as f:
        f.write(_stamp())
    return LIB


if __name__ == "__main__":
    print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
