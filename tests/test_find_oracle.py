"""CPU: the oracle's restatement of find / find_idx (oracle/matx_oracle.c: orc_find, following find_impl / find_idx_impl
for HostExecutor, transforms/cub.h:2656-2675,2752-2770) pinned against (a) the known-answer bodies of the reference's
tests (test/00_operators/ReductionTests.cu:1615-1693: every t1(i) > thresh in order, count == num_found) and (b) golden
vectors produced by THE REFERENCE ITSELF (tests/golden/reference_find.npz, made by tests/golden/make_golden_find.py from
oracle/_ref), plus live calls of that library when it is present."""
import os

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import oracle_harness as H
from tests.oracle_harness import np_tensor

SEL = [mx.LT, mx.GT, mx.EQ, mx.NEQ, mx.LTE, mx.GTE]
NPSEL = [np.less, np.greater, np.equal, np.not_equal, np.less_equal, np.greater_equal]
GOLD = os.path.join(os.path.dirname(__file__), "golden", "reference_find.npz")


def test_reference_test_bodies(oracle):
    # ReductionTests.cu:1624-1647 (Find) and :1665-1688 (FindIdx): 100 random values in [0, 2), thresh 0.5, GT
    rng = np.random.default_rng(1)
    t1 = (rng.random(100) * 2).astype(np.float32)
    out = np.zeros(100, np.float32)
    n = oracle.find(mx.find(np_tensor(t1), mx.GT(0.5)), out)
    want = t1[t1 > 0.5]
    assert n == len(want) and np.array_equal(out[:n], want)
    idx = np.zeros(100, np.int32)
    n2 = oracle.find(mx.find_idx(np_tensor(t1), mx.GT(0.5)), idx)
    assert n2 == n and np.array_equal(idx[:n], np.nonzero(t1 > 0.5)[0])
    # FindIdxAndSelect (:1695-1737): selecting by the found indices gives the found values
    assert np.array_equal(t1[idx[:n]], out[:n])


@pytest.mark.parametrize("sel", range(6))
def test_all_functors_views_and_dtypes(oracle, sel):
    rng = np.random.default_rng(2 + sel)
    x = (rng.integers(0, 9, (23, 31)) * 0.25).astype(np.float32)
    for view in (x, x[3:20, 2:29], x.T, x.reshape(-1)[5::4]):
        want = view.reshape(-1)[NPSEL[sel](view.reshape(-1), np.float32(1.0))] if view.flags.c_contiguous else \
            np.ascontiguousarray(view).reshape(-1)[NPSEL[sel](np.ascontiguousarray(view).reshape(-1), np.float32(1.0))]
        out = np.zeros(view.size, np.float32)
        n = oracle.find(mx.find(np_tensor(view), SEL[sel](1.0)), out)
        assert n == len(want) and np.array_equal(out[:n], want)
        idx = np.zeros(view.size, np.int64)
        n = oracle.find(mx.find_idx(np_tensor(view), SEL[sel](1.0)), idx)
        assert np.array_equal(idx[:n], np.nonzero(NPSEL[sel](np.ascontiguousarray(view).reshape(-1), np.float32(1.0)))[0])
    xi = rng.integers(-5, 6, 400).astype(np.int32)
    out = np.zeros(400, np.int32)
    n = oracle.find(mx.find(np_tensor(xi), SEL[sel](2)), out)
    assert np.array_equal(out[:n], xi[NPSEL[sel](xi, 2)])
    # an expression as the operand, a capacity smaller than the count, an empty selection
    e = np_tensor(x) * 2.0 - 1.0
    ev = (x * np.float32(2.0) - np.float32(1.0)).reshape(-1)
    small = np.zeros(7, np.float32)
    n = oracle.find(mx.find(e, SEL[sel](0.5)), small)
    want = ev[NPSEL[sel](ev, np.float32(0.5))]
    assert n == len(want) and np.array_equal(small[:min(n, 7)], want[:7])
    n = oracle.find(mx.find(np_tensor(x), mx.GT(100.0)), out.astype(np.float32))
    assert n == 0


def test_oracle_reproduces_reference_golden_vectors(oracle):
    g = np.load(GOLD)
    tags = sorted({k.split("/")[0] for k in g.files})
    n_cases = 0
    for tag in tags:
        xs = g[tag + "/x"]
        for sel in range(6):
            for thr in (0.5, 1.0, 9.0):
                for want_idx in (0, 1):
                    key = "%s/sel%d/thr%g/idx%d" % (tag, sel, thr, want_idx)
                    want, wn = g[key + "/out"], int(g[key + "/n"])
                    out = np.zeros(xs.size, np.int32 if want_idx else np.float32)
                    t = np_tensor(xs)   # the flat order of a view is the order of its (contiguous) copy
                    r = mx.find_idx(t, SEL[sel](thr)) if want_idx else mx.find(t, SEL[sel](thr))
                    n = oracle.find(r, out)
                    assert n == wn and np.array_equal(out[:n], want), key
                    n_cases += 1
    assert n_cases >= 200


def test_oracle_vs_live_reference(oracle):
    ref = H.load_ref_host()
    if ref is None or not hasattr(ref.lib, "mref_find_f32_host"):
        pytest.skip("oracle/_ref is not built here (the reference is not on this box)")
    from tests.golden.make_golden_find import ref_find
    rng = np.random.default_rng(77)
    x = (rng.integers(0, 5, (40, 64)) * 0.5).astype(np.float32)
    for view in (x, x[1:33, 7:60], x.T):
        for sel in range(6):
            for want_idx in (False, True):
                want, wn = ref_find(ref, view, sel, 1.0, want_idx, view.size)
                out = np.zeros(view.size, np.int32 if want_idx else np.float32)
                r = mx.find_idx(np_tensor(view), SEL[sel](1.0)) if want_idx else mx.find(np_tensor(view), SEL[sel](1.0))
                n = oracle.find(r, out)
                assert n == wn and np.array_equal(out[:n], want[:wn]), (sel, want_idx)
