"""Build libmatx_b200.so in-tree (nvcc cross-compiles sm_100a without a GPU).

Steps: (1) g++ builds the generator `mxb_gen` from codegen.cpp + aot_manifest.cpp, (2) mxb_gen prints the
ahead-of-time kernel shards and the embedded skeleton header into csrc/_gen/, (3) nvcc compiles the shards
and api.cu for sm_100a, g++ the host files, (4) one link into matx_b200/libmatx_b200.so.

Run as `python -m matx_b200.build [--force]`; `__graft_entry__.build()` calls `build()`.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
GEN = os.path.join(CSRC, "_gen")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libmatx_b200.so")
SHARDS = 8

NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
    "--expt-relaxed-constexpr",
]
CXX_FLAGS = ["-std=c++17", "-O2", "-fPIC"]


def _cuda_include() -> str:
    """<cuda.h> for the host files that use driver TYPES (entry points are resolved at run time with dlsym)."""
    return os.path.join(os.path.dirname(os.path.dirname(os.path.realpath(_nvcc()))), "include")

HOST_SOURCES = ["codegen.cpp", "jit.cpp", "aot_manifest.cpp"]
GEN_SOURCES = ["gen_main.cpp", "codegen.cpp", "aot_manifest.cpp"]
ALL_INPUTS = ["mxb_device.cuh", "mxb_internal.h", "mxb_sort.cuh", "api.cu", "gen_main.cpp"] + HOST_SOURCES


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (needed to build libmatx_b200.so)")


def _digest() -> str:
    h = hashlib.sha256()
    for name in ALL_INPUTS + [os.path.join("..", "..", "include", "matx_b200.h"), os.path.join("..", "build.py")]:
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(name.encode())
            h.update(f.read())
    return h.hexdigest()


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout))


def build(force: bool = False, verbose: bool = False) -> str:
    stamp = os.path.join(OBJ, "stamp")
    digest = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == digest:
        return LIB
    os.makedirs(GEN, exist_ok=True)
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    gen = os.path.join(OBJ, "mxb_gen")
    _run(["g++", "-std=c++17", "-O1", "-o", gen] + [os.path.join(CSRC, s) for s in GEN_SOURCES])
    for f in os.listdir(GEN):
        os.remove(os.path.join(GEN, f))
    _run([gen, CSRC, GEN, str(SHARDS)])

    jobs: list[tuple[list[str], str]] = []
    objs: list[str] = []
    for i in range(SHARDS):
        o = os.path.join(OBJ, "aot_%d.o" % i)
        objs.append(o)
        jobs.append(([nvcc] + NVCC_FLAGS + ["-c", os.path.join(GEN, "aot_%d.cu" % i), "-o", o], o))
    o = os.path.join(OBJ, "api.o")
    objs.append(o)
    jobs.append(([nvcc] + NVCC_FLAGS + ["-c", os.path.join(CSRC, "api.cu"), "-o", o], o))
    for s in HOST_SOURCES:
        o = os.path.join(OBJ, s.replace(".cpp", ".o"))
        objs.append(o)
        jobs.append((["g++"] + CXX_FLAGS + ["-I" + _cuda_include(), "-c", os.path.join(CSRC, s), "-o", o], o))
    o = os.path.join(OBJ, "device_src.o")
    objs.append(o)
    jobs.append((["g++"] + CXX_FLAGS + ["-c", os.path.join(GEN, "device_src.cpp"), "-o", o], o))

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 2)) as ex:
        for fut in [ex.submit(_run, cmd) for cmd, _ in jobs]:
            fut.result()
    _run([nvcc, "-shared", "-o", LIB] + objs + ["-ldl", "-lpthread"])
    with open(stamp, "w") as f:
        f.write(digest)
    if verbose:
        print("built", LIB)
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose=True)
