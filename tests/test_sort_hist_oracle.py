"""CPU: the oracle's sort / unique / hist restatements (oracle/matx_oracle.c) against (a) the reference's own known answers
(test/00_tensor/CUBTests.cu:153-197 hist, test/00_operators/ReductionTests.cu:1744-1770 unique) and (b) golden vectors the
reference itself produced here (tests/golden/reference_sort.npz, made by tests/golden/make_golden_sort.py from
oracle/_ref/libmatx_ref_host.so).  hist has no HostExecutor form in the reference: it is pinned by (a) only — "pinned to
third party" (CCCL's HistogramEven), as DESIGN.md section 5 says."""
import os

import numpy as np
import pytest

from matx_b200 import ops as mx
from tests.oracle_harness import np_tensor

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_sort.npz")


def test_hist_known_answers_of_the_reference(oracle):
    x = np.array([2.2, 6.0, 7.1, 2.9, 3.5, 0.3, 2.9, 2.0, 6.1, 999.5], np.float32)       # CUBTests.cu:153-176
    out = np.full(6, -1, np.int32)
    oracle.hist(mx.hist(np_tensor(x), 0.0, 12.0, 7), out)
    assert out.tolist() == [1, 5, 0, 3, 0, 0]
    s = np.array([0, 99, 1, 99, 2, 99, 0, 99, 1, 99, 2, 99], np.float32)                    # CUBTests.cu:178-197 (strided input)
    out = np.full(3, -1, np.int32)
    oracle.hist(mx.hist(np_tensor(s[::2]), 0.0, 3.0, 4), out)
    assert out.tolist() == [2, 2, 2]
    # bounds: lower is inside, upper is outside; integer samples use the integer formula
    xi = np.array([0, 1, 9, 10, 5, 5, -1], np.int32)
    out = np.full(5, -1, np.int32)
    oracle.hist(mx.hist(np_tensor(xi), 0, 10, 6), out)
    assert out.tolist() == [2, 0, 2, 0, 1]


def test_unique_known_answer_of_the_reference(oracle):
    t = (np.arange(100) % 10).astype(np.float32)                                             # ReductionTests.cu:1744-1770
    out = np.full(100, -7, np.float32)
    n = oracle.unique(mx.unique(np_tensor(t)), out)
    assert n == 10 and out[:10].tolist() == list(range(10))


def test_sort_and_unique_match_the_golden_vectors_of_the_reference(oracle):
    if not os.path.exists(GOLD):
        pytest.skip("tests/golden/reference_sort.npz missing")
    g = np.load(GOLD)
    tags = sorted({k.split("/")[0] for k in g.files})
    assert len(tags) >= 6
    for tag in tags:
        x = g[tag + "/x"]
        for name, direction in (("asc", mx.SORT_DIR_ASC), ("desc", mx.SORT_DIR_DESC)):
            out = np.zeros_like(x)
            oracle.sort(mx.sort(np_tensor(x), direction), out)
            assert np.array_equal(out, g[tag + "/" + name]), (tag, name)
        if tag + "/unique" in g.files:
            out = np.zeros_like(x)
            n = oracle.unique(mx.unique(np_tensor(x)), out)
            assert n == int(g[tag + "/unique_n"]) and np.array_equal(out[:n], g[tag + "/unique"]), tag


def test_sort_of_an_expression_and_a_strided_view(oracle):
    rng = np.random.default_rng(3)
    a = rng.standard_normal((8, 50)).astype(np.float32)
    out = np.zeros((8, 25), np.float32)
    oracle.sort(mx.sort(np_tensor(a[:, ::2]) * 2.0, mx.SORT_DIR_ASC), out)
    assert np.array_equal(out, np.sort(a[:, ::2] * np.float32(2.0), axis=1))
