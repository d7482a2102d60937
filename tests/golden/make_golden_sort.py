"""Golden vectors for sort / unique from THE REFERENCE ITSELF: oracle/_ref/libmatx_ref_host.so (matx::HostExecutor
statements compiled from /root/reference by oracle/build_ref.py) -> tests/golden/reference_sort.npz.
Run here (the reference is not on the GPU box):  python tests/golden/make_golden_sort.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_harness as H  # noqa: E402

SUFFIX = {np.dtype(np.float32): "f32", np.dtype(np.float64): "f64", np.dtype(np.int32): "i32"}


def ref_sort(ref, x: np.ndarray, desc: bool) -> np.ndarray:
    f = ref.fn("mref_sort_" + SUFFIX[x.dtype])
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    x = np.ascontiguousarray(x)
    rows, cols = (1, x.shape[0]) if x.ndim == 1 else x.shape
    out = np.zeros_like(x)
    if f(0, 1 if desc else 0, x.ctypes.data, out.ctypes.data, rows, cols) != 0:
        raise RuntimeError("reference sort failed")
    return out


def ref_unique(ref, x: np.ndarray):
    f = ref.fn("mref_unique_" + SUFFIX[x.dtype])
    f.argtypes = [C.c_int, C.c_void_p, C.c_int64, C.c_void_p, C.POINTER(C.c_int)]
    x = np.ascontiguousarray(x)
    out = np.zeros_like(x)
    n = C.c_int(-1)
    if f(0, x.ctypes.data, x.size, out.ctypes.data, C.byref(n)) != 0:
        raise RuntimeError("reference unique failed")
    return out[: n.value].copy(), int(n.value)


def cases():
    rng = np.random.default_rng(77)
    yield "f32_1d", (rng.standard_normal(5000) * 100).astype(np.float32)
    yield "f32_ties", (rng.integers(-20, 20, 3000) * 0.5).astype(np.float32)
    yield "f32_rows", (rng.standard_normal((37, 129))).astype(np.float32)
    yield "f64_1d", rng.standard_normal(2049) * 1e3
    yield "i32_1d", rng.integers(-1000, 1000, 4097).astype(np.int32)
    yield "i32_rows", rng.integers(-5, 5, (16, 300)).astype(np.int32)
    yield "f32_cubtests", np.array([-1.0, 2.5, 7.0, -3.25, 0.5, 7.0, 1e6, -1e-6, 0.0, 3.0], np.float32)


def main():
    ref = H.load_ref_host()
    if ref is None:
        raise SystemExit("oracle/_ref/libmatx_ref_host.so is missing: python oracle/build_ref.py")
    out = {}
    for tag, x in cases():
        out[tag + "/x"] = x
        out[tag + "/asc"] = ref_sort(ref, x, False)
        out[tag + "/desc"] = ref_sort(ref, x, True)
        if x.ndim == 1 and x.dtype != np.float64:
            u, n = ref_unique(ref, x)
            out[tag + "/unique"] = u
            out[tag + "/unique_n"] = np.int32(n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_sort.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
