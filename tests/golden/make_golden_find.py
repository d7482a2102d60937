"""Golden vectors for find / find_idx from THE REFERENCE ITSELF: oracle/_ref/libmatx_ref_host.so (matx::HostExecutor
statements compiled from /root/reference by oracle/build_ref.py) -> tests/golden/reference_find.npz.
Run here (the reference is not on the GPU box):  python tests/golden/make_golden_find.py"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tests import oracle_harness as H  # noqa: E402


def ref_find(ref, x: np.ndarray, sel: int, thr: float, want_idx: bool, cap: int):
    f = ref.fn("mref_find_f32")
    f.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_void_p, C.c_int64,
                  C.POINTER(C.c_int), C.c_int]
    rank = x.ndim
    sh = (C.c_int64 * rank)(*x.shape)
    st = (C.c_int64 * rank)(*[s // 4 for s in x.strides])
    out = np.full(cap, -7, np.int32 if want_idx else np.float32)
    n = C.c_int(-1)
    rc = f(0, sel, thr, rank, sh, st, C.c_void_p(x.ctypes.data), C.c_void_p(out.ctypes.data), cap, C.byref(n), 1 if want_idx else 0)
    if rc != 0:
        raise RuntimeError("reference find failed")
    return out, int(n.value)


def cases():
    rng = np.random.default_rng(2024)
    base = (rng.integers(0, 9, (37, 53)) * 0.25).astype(np.float32)          # many ties with the thresholds
    yield "r1", base.reshape(-1)[:1500].copy()
    yield "r1_strided", base.reshape(-1)[3:1800:3]
    yield "r2", base
    yield "r2_sliced", base[2:30, 5:47]
    yield "r2_transposed", base.T
    yield "r1_short", np.array([0.5, 0.75, 0.5, 0.25, 1.0], np.float32)


def main():
    ref = H.load_ref_host()
    if ref is None:
        raise SystemExit("oracle/_ref/libmatx_ref_host.so is missing: python oracle/build_ref.py")
    out = {}
    for tag, x in cases():
        out[tag + "/x"] = np.ascontiguousarray(x)
        out[tag + "/strides"] = np.array([s // 4 for s in x.strides], np.int64)
        for sel in range(6):
            for thr in (0.5, 1.0, 9.0):
                for want_idx in (0, 1):
                    vals, n = ref_find(ref, x, sel, thr, bool(want_idx), x.size)
                    out["%s/sel%d/thr%g/idx%d/out" % (tag, sel, thr, want_idx)] = vals[:n].copy()
                    out["%s/sel%d/thr%g/idx%d/n" % (tag, sel, thr, want_idx)] = np.int32(n)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "reference_find.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
