"""Generate tests/golden/reference_host.npz from THE REFERENCE ITSELF: matx::HostExecutor<SINGLE> statements compiled from
/root/reference (oracle/_ref/libmatx_ref_host.so, see oracle/ref_wrap.cu), run in this container on seeded inputs.
The reference cannot travel to the GPU box as source, so the input/output vectors are committed; this script is the
record of how they were made.

    python oracle/build_ref.py && python tests/golden/make_golden.py
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from matx_b200 import _abi as A  # noqa: E402
from tests import oracle_harness as H  # noqa: E402

OPS = {"sum": A.RED_SUM, "mean": A.RED_MEAN, "var": A.RED_VAR, "stdd": A.RED_STDD, "max": A.RED_MAX, "min": A.RED_MIN,
       "argmax": A.RED_ARGMAX, "argmin": A.RED_ARGMIN, "any": A.RED_ANY, "all": A.RED_ALL, "prod": A.RED_PROD}

# (name, dtype, shape, permutation applied to the contiguous array before reducing, dims reduced)
CASES = [
    ("f32_1d", np.float32, (1000,), None, [0]),
    ("f32_rows", np.float32, (37, 129), None, [1]),
    ("f32_cols", np.float32, (37, 129), None, [0]),
    ("f32_3d_inner2", np.float32, (6, 33, 33), None, [1, 2]),
    ("f32_3d_mid", np.float32, (6, 33, 33), None, [1]),
    ("f32_4d_perm", np.float32, (12, 10, 9, 16), [2, 3, 0, 1], [2, 3]),   # PermutedReduce, ReductionTests.cu:353-573
    ("f32_4d_01", np.float32, (12, 10, 9, 16), None, [0, 1]),
    ("f32_full_4d", np.float32, (5, 6, 7, 8), None, [0, 1, 2, 3]),
    ("f64_rows", np.float64, (20, 300), None, [1]),
    ("f64_full", np.float64, (3000,), None, [0]),
    ("i32_rows", np.int32, (25, 100), None, [1]),
    ("i32_cols", np.int32, (25, 100), None, [0]),
    ("c64_rows", np.complex64, (16, 512), None, [1]),
    ("c64_full", np.complex64, (700,), None, [0]),
]


def make_input(rng, dt, shape, ties):
    if dt == np.complex64:
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape) + (1 + 0.5j)).astype(dt)
    if dt == np.int32:
        return rng.integers(-9, 10, shape).astype(dt)
    if ties:
        return rng.integers(0, 6, shape).astype(dt)
    return (rng.random(shape) + 0.25).astype(dt)


def main():
    ref = H.load_ref_host()
    assert ref is not None, "build oracle/_ref first"
    out = {}
    rng = np.random.default_rng(20261017)
    for name, dt, shape, perm, dims in CASES:
        for ties in (False, True):
            if ties and dt in (np.complex64, np.int32):
                continue
            x = make_input(rng, dt, shape, ties)
            tag = name + ("_ties" if ties else "")
            out[tag + "/x"] = x
            view = np.transpose(x, perm) if perm else x
            for opn, op in OPS.items():
                if dt == np.complex64 and opn in ("max", "min", "argmax", "argmin"):
                    continue
                if dt == np.int32 and opn in ("var", "stdd", "mean"):
                    continue
                if opn == "prod":
                    n_red = int(np.prod([view.shape[d] for d in dims]))
                    if n_red > 200:
                        continue
                v, i = ref.reduce_np(op, view, dims, ddof=1, mode=0)
                out["%s/%s" % (tag, opn)] = v
                if i is not None:
                    out["%s/%s_idx" % (tag, opn)] = i
    # fused statements
    a, b = (rng.random((48, 1024)).astype(np.float32) for _ in range(2))
    c = (rng.random((48, 1024)) - 0.5).astype(np.float32)
    o = np.zeros(48, np.float32)
    p = lambda z: z.ctypes.data_as(C.c_void_p)  # noqa: E731
    f = ref.fn("mref_fma_sum")
    f.argtypes = [C.c_int] + [C.c_void_p] * 4 + [C.c_int64, C.c_int64]
    assert f(0, p(a), p(b), p(c), p(o), 48, 1024) == 0
    out.update({"fma_sum/a": a, "fma_sum/b": b, "fma_sum/c": c, "fma_sum/out": o})
    x = (rng.standard_normal((32, 512)) + 1j * rng.standard_normal((32, 512))).astype(np.complex64)
    v, ix = np.zeros(32, np.float32), np.zeros(32, np.int64)
    f = ref.fn("mref_abs2_argmax")
    f.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int64, C.c_int64]
    assert f(0, p(x), p(v), p(ix), 32, 512) == 0
    out.update({"abs2_argmax/x": x, "abs2_argmax/val": v, "abs2_argmax/idx": ix})
    n = 4096
    S, K = (rng.uniform(10, 100, n).astype(np.float32) for _ in range(2))
    V = rng.uniform(0.05, 0.5, n).astype(np.float32)
    r = rng.uniform(0.01, 0.1, n).astype(np.float32)
    T = rng.uniform(0.1, 2, n).astype(np.float32)
    price = np.zeros(n, np.float32)
    f = ref.fn("mref_black_scholes")
    f.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_int64]
    assert f(0, p(K), p(S), p(V), p(r), p(T), p(price), n) == 0
    out.update({"bs/K": K, "bs/S": S, "bs/V": V, "bs/r": r, "bs/T": T, "bs/out": price})
    xa = (rng.random(512) * 0.98 + 0.01).astype(np.float32)
    xb = (rng.random(512) + 0.5).astype(np.float32)
    out["functor/a"], out["functor/b"] = xa, xb
    f = ref.fn("mref_unary_f32")
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
    for opc in [40, 41, 42, 43, 44, 45, 46, 47, 48, 52, 53, 54, 55, 56, 60, 61, 62, 63, 64, 65, 66, 67]:
        y = np.zeros(512, np.float32)
        assert f(0, opc, p(xa), p(y), 512) == 0, opc
        out["functor/unary_%d" % opc] = y
    f = ref.fn("mref_binary_f32")
    f.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64]
    for opc in [10, 11, 12, 13, 14, 15, 16, 17]:
        y = np.zeros(512, np.float32)
        assert f(0, opc, p(xa), p(xb), p(y), 512) == 0, opc
        out["functor/binary_%d" % opc] = y
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_host.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
