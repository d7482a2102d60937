"""Shared description of tests/golden/reference_host.npz (made by tests/golden/make_golden.py from the reference's
HostExecutor): which statement produced which array."""
import os

import numpy as np

from matx_b200 import ops as mx

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_host.npz")

# (name, permutation applied to the stored contiguous input, dims reduced) — keep in step with make_golden.py
CASES = {
    "f32_1d": (None, [0]), "f32_rows": (None, [1]), "f32_cols": (None, [0]), "f32_3d_inner2": (None, [1, 2]),
    "f32_3d_mid": (None, [1]), "f32_4d_perm": ([2, 3, 0, 1], [2, 3]), "f32_4d_01": (None, [0, 1]),
    "f32_full_4d": (None, [0, 1, 2, 3]), "f64_rows": (None, [1]), "f64_full": (None, [0]), "i32_rows": (None, [1]),
    "i32_cols": (None, [0]), "c64_rows": (None, [1]), "c64_full": (None, [0]),
}
EXACT = ("max", "min", "argmax", "argmin", "any", "all")


def load():
    return np.load(PATH)


def statements(g):
    """Yield (tag, opname, build(tensor) -> ReduceExpr, golden values, golden indices or None)."""
    for key in g.files:
        if not key.endswith("/x"):
            continue
        tag = key[:-2]
        base = tag[:-5] if tag.endswith("_ties") else tag
        if base not in CASES:
            continue  # inputs of the fused statements
        perm, dims = CASES[base]
        for opn in ("sum", "mean", "var", "stdd", "max", "min", "argmax", "argmin", "any", "all", "prod"):
            k = "%s/%s" % (tag, opn)
            if k not in g.files:
                continue

            def build(t, opn=opn, perm=perm, dims=dims):
                v = mx.permute(t, perm) if perm else t
                f = getattr(mx, opn)
                return f(v, dims, 1) if opn in ("var", "stdd") else f(v, dims)

            yield tag, opn, build, g[k], (g[k + "_idx"] if (k + "_idx") in g.files else None)
