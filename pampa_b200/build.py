"""In-tree build of the native libraries (no JIT cache: the .so files travel with the tree).

  pampa_b200/lib/libpampa_sn_b200.so   CUDA layer + C ABI of include/pampa_sn.h   (nvcc, sm_100a)
  pampa_b200/lib/libpampa.so           C++ host code + C API of include/pampa.h   (g++)
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "lib")
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _run(cmd):
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("build failed: " + " ".join(cmd))


def _sources(d, exts):
    return sorted(os.path.join(d, f) for f in os.listdir(d) if f.endswith(exts)) if os.path.isdir(d) else []


def build_cuda(force=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libpampa_sn_b200.so")
    cu = _sources(CSRC, (".cu",))
    deps = cu + _sources(CSRC, (".cuh", ".hpp")) + _sources(os.path.join(ROOT, "include"), (".h",))
    if force or _newer(out, deps):
        _run(["nvcc"] + NVCC_FLAGS + ["--threads", str(max(1, len(cu)))] + cu + ["-o", out, "-ldl"])
    return out


def build_host(force=False):
    os.makedirs(LIB, exist_ok=True)
    out = os.path.join(LIB, "libpampa.so")
    cpp = _sources(HOST, (".cpp",))
    if not cpp:
        return None
    deps = cpp + _sources(HOST, (".hpp",)) + _sources(os.path.join(ROOT, "include"), (".h",))
    if force or _newer(out, deps):
        _run(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-Wall"] + cpp +
             ["-o", out, "-L" + LIB, "-lpampa_sn_b200", "-Wl,-rpath,$ORIGIN"])
    return out


def build_main(force=False):
    """bin/pampa: the stand-alone driver (analogue of the reference's cxx main)."""
    src = os.path.join(HOST, "main.cxx")
    out = os.path.join(HERE, "bin", "pampa")
    if not os.path.exists(src):
        return None
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if force or _newer(out, [src, os.path.join(LIB, "libpampa.so")]):
        _run(["g++", "-O2", "-std=c++17", src, "-o", out, "-L" + LIB, "-lpampa", "-lpampa_sn_b200",
              "-Wl,-rpath,$ORIGIN/../lib"])
    return out


def build_all(force=False):
    return [build_cuda(force), build_host(force), build_main(force)]


if __name__ == "__main__":
    for p in build_all(force="--force" in sys.argv):
        print(p)
