"""Build the checkers under oracle/:

  oracle/_build/libmatx_oracle.so   gcc, from oracle/matx_oracle.c (the CPU restatement; always buildable)
  oracle/_ref/libmatx_ref_host.so   nvcc, reference HostExecutor statements (ref_wrap.cu) — needs /root/reference
  oracle/_ref/libmatx_ref_cuda.so   nvcc, reference cudaExecutor + CUB statements   — needs /root/reference

The reference is compiled from its headers WHERE THEY LIE (/root/reference/include); nothing is copied.  The
reference pins CCCL 3.3.0 (cmake/versions.json:3-8) which is not vendored; the toolkit's CCCL 2.8.2 is too old,
so the CCCL 3.3.2 copy shipped inside the flashinfer wheel is used.  The reference's own
build system (CMake + CPM, needs network) is not run.  Outputs under oracle/_ref/ are git-ignored but travel to
the GPU box with the snapshot; on the box /root/reference does not exist and the prebuilt files are used as-is.

    python oracle/build_ref.py [--cuda] [--force]
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("MATX_REFERENCE", "/root/reference")
CCCL = os.environ.get("MATX_CCCL", "/opt/prime-rl/.venv/lib/python3.12/site-packages/flashinfer/data/cccl")
OUT_REF = os.path.join(HERE, "_ref")
OUT_BUILD = os.path.join(HERE, "_build")
ORACLE_LIB = os.path.join(OUT_BUILD, "libmatx_oracle.so")
REF_HOST_LIB = os.path.join(OUT_REF, "libmatx_ref_host.so")
REF_CUDA_LIB = os.path.join(OUT_REF, "libmatx_ref_cuda.so")

# -march=x86-64-v3 (AVX2/FMA-capable, runs on any current server) instead of -march=native: the library is
# built here and executed on the GPU box's CPU.  -ffp-contract=off keeps a*b+c as two roundings, the same
# arithmetic the C restatement performs, so the two can be compared bit for bit.
HOST_CXX = "-fopenmp,-fPIC,-O3,-march=x86-64-v3,-ffp-contract=off"


def _run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        raise RuntimeError("command failed: %s\n%s" % (" ".join(cmd), r.stdout[-6000:]))


def build_oracle(force: bool = False) -> str:
    src = os.path.join(HERE, "matx_oracle.c")
    if not force and os.path.exists(ORACLE_LIB) and os.path.getmtime(ORACLE_LIB) >= os.path.getmtime(src):
        return ORACLE_LIB
    os.makedirs(OUT_BUILD, exist_ok=True)
    _run(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", ORACLE_LIB, src, "-lm"])
    return ORACLE_LIB


def reference_available() -> bool:
    return os.path.exists(os.path.join(REF, "include", "matx.h")) and os.path.isdir(CCCL)


def _nvcc_base(cuda: bool) -> list[str]:
    cmd = ["nvcc", "-std=c++20", "-gencode", "arch=compute_100a,code=sm_100a", "--extended-lambda", "-O3", "-w",
           "-Xcompiler", HOST_CXX, "-DMATX_EN_OMP", "-DNDEBUG",
           "-I" + os.path.join(CCCL, "libcudacxx", "include"), "-I" + os.path.join(CCCL, "cub"),
           "-I" + os.path.join(CCCL, "thrust"), "-I" + os.path.join(REF, "include"),
           "-I" + os.path.join(REF, "include", "matx", "kernels")]
    if cuda:
        cmd.append("-DMREF_CUDA")
    return cmd


def build_ref(cuda: bool = False, force: bool = False) -> str | None:
    """Compile the reference statements.  Returns the library path, or None when /root/reference is absent."""
    lib = REF_CUDA_LIB if cuda else REF_HOST_LIB
    src = os.path.join(HERE, "ref_wrap.cu")
    if not reference_available():
        return lib if os.path.exists(lib) else None
    if not force and os.path.exists(lib) and os.path.getmtime(lib) >= os.path.getmtime(src):
        return lib
    os.makedirs(OUT_REF, exist_ok=True)
    tag = "cuda" if cuda else "host"
    # one translation unit per element type (the reduce instantiation sets are large) plus the fused statements
    units = [("dt0", ["-DMREF_DTYPE=0"]), ("dt4", ["-DMREF_DTYPE=4"]), ("dt2", ["-DMREF_DTYPE=2"]), ("fused", ["-DMREF_FUSED"]),
             ("sort", ["-DMREF_SORT"])]
    if not cuda:
        units += [("dt1", ["-DMREF_DTYPE=1"]), ("dt5", ["-DMREF_DTYPE=5"])]
    objs = []
    jobs = []
    for name, defs in units:
        o = os.path.join(OUT_REF, "ref_%s_%s.o" % (tag, name))
        objs.append(o)
        if os.environ.get("MREF_REUSE_OBJS") and os.path.exists(o):
            continue  # development shortcut: keep objects of units whose part of ref_wrap.cu did not change
        jobs.append(_nvcc_base(cuda) + defs + ["-c", src, "-o", o])
    with cf.ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 2))) as ex:
        for fut in [ex.submit(_run, j) for j in jobs]:
            fut.result()
    _run(["nvcc", "-shared", "-o", lib] + objs + ["-Xcompiler", "-fopenmp", "-lgomp"])
    for o in objs:
        os.remove(o)
    return lib


DROPIN_BIN = os.path.join(OUT_REF, "dropin_test")


def build_dropin(force: bool = False) -> str | None:
    """tests/cpp/dropin_test.cu: unmodified MatX statements on matx::cudaExecutor vs matx::b200Executor (the header
    shim include/matx_b200/executor.h over libmatx_b200.so).  Needs /root/reference; the binary travels to the GPU box."""
    root = os.path.dirname(HERE)
    src = os.path.join(root, "tests", "cpp", "dropin_test.cu")
    shim = os.path.join(root, "include", "matx_b200", "executor.h")
    lib = os.path.join(root, "matx_b200", "libmatx_b200.so")
    if not reference_available():
        return DROPIN_BIN if os.path.exists(DROPIN_BIN) else None
    if not os.path.exists(lib):
        raise RuntimeError("build matx_b200/libmatx_b200.so first (python -m matx_b200.build)")
    newest = max(os.path.getmtime(src), os.path.getmtime(shim), os.path.getmtime(os.path.join(root, "include", "matx_b200.h")))
    if not force and os.path.exists(DROPIN_BIN) and os.path.getmtime(DROPIN_BIN) >= newest:
        return DROPIN_BIN
    os.makedirs(OUT_REF, exist_ok=True)
    # overlay headers (operators/softmax.h, operators/hist.h handing their impl the executor): generated from the reference's
    # own text into oracle/_ref/overlay and put in FRONT of the reference's include directory
    sys.path.insert(0, os.path.join(root, "tools"))
    import make_overlay
    overlay = make_overlay.main()
    base = _nvcc_base(False)
    ref_inc = "-I" + os.path.join(REF, "include")
    base.insert(base.index(ref_inc), "-I" + overlay)
    cmd = base + ["-DMXB_OVERLAY", "-I" + os.path.join(root, "include"), src, "-o", DROPIN_BIN, "-L" + os.path.join(root, "matx_b200"),
                               "-lmatx_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../../matx_b200",
                               "-lcublas", "-lcublasLt", "-lcufft", "-lcurand", "-lcusolver", "-lcusparse", "-lcuda"]
    _run(cmd)
    return DROPIN_BIN


if __name__ == "__main__":
    force = "--force" in sys.argv
    print(build_oracle(force))
    print(build_ref(False, force))
    if "--cuda" in sys.argv:
        print(build_ref(True, force))
    if "--dropin" in sys.argv:
        print(build_dropin(force))
