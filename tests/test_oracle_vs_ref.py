"""Pins oracle/matx_oracle.c against OUTPUTS OF THE REFERENCE ITSELF: the golden vectors made by running
matx::HostExecutor (compiled from /root/reference) — tests/golden/reference_host.npz — and, when oracle/_ref is
present, live calls of the same library on fresh random inputs.  CPU only.

fp32 / fp64 / int sums, means, products, var / stdd and every exact op must agree BIT FOR BIT: the oracle restates
the host path's sequential value-type arithmetic, and both are built without FMA contraction."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import golden_util as GU
from tests import oracle_harness as H
from tests.oracle_harness import np_tensor

NP2 = {np.dtype(np.float32): A.F32, np.dtype(np.float64): A.F64, np.dtype(np.complex64): A.C64, np.dtype(np.int32): A.I32}


def test_oracle_reproduces_reference_golden_vectors(oracle):
    g = GU.load()
    n = 0
    for tag, opn, build, want, widx in GU.statements(g):
        x = g[tag + "/x"]
        r = build(np_tensor(x))
        out = np.zeros(r.out_shape, want.dtype)
        idx = np.zeros(r.out_shape, np.int64) if widx is not None else None
        oracle.reduce(r, out, idx)
        assert np.array_equal(out, want, equal_nan=True), (tag, opn, out.ravel()[:4], want.ravel()[:4])
        if widx is not None:
            assert np.array_equal(idx, widx), (tag, opn)
        n += 1
    assert n > 150


def test_oracle_reproduces_reference_fused_statements(oracle):
    g = GU.load()
    a, b, c = (np_tensor(g["fma_sum/" + k]) for k in "abc")
    out = np.zeros(48, np.float32)
    oracle.reduce(mx.sum(a * b + c, [1]), out)
    assert np.array_equal(out, g["fma_sum/out"])
    x = np_tensor(g["abs2_argmax/x"])
    v, i = np.zeros(32, np.float32), np.zeros(32, np.int64)
    oracle.reduce(mx.argmax(mx.abs2(x), [1]), v, i)
    assert np.array_equal(v, g["abs2_argmax/val"]) and np.array_equal(i, g["abs2_argmax/idx"])


def test_oracle_functors_match_reference(oracle):
    g = GU.load()
    a, b = g["functor/a"], g["functor/b"]
    names = {40: "neg", 41: "sqrt", 42: "rsqrt", 43: "exp", 44: "log", 45: "log2", 46: "log10", 47: "abs", 48: "abs2", 52: "sin",
             53: "cos", 54: "tan", 55: "tanh", 56: "normcdf", 60: "floor", 61: "ceil", 62: "round_", 63: "sinh", 64: "cosh",
             65: "asin", 66: "acos", 67: "atan"}
    for opc, nm in names.items():
        out = np.zeros(512, np.float32)
        rhs = -np_tensor(a) if nm == "neg" else getattr(mx, nm)(np_tensor(a))
        oracle.elementwise(rhs, out)
        want = g["functor/unary_%d" % opc]
        # same libm underneath except normcdf (CUDA's host emulation vs erfc): a few ulp
        assert np.allclose(out, want, rtol=4e-7, atol=1e-7), (nm, np.max(np.abs(out - want)))
    bins = {10: lambda x, y: x + y, 11: lambda x, y: x - y, 12: lambda x, y: x * y, 13: lambda x, y: x / y,
            14: lambda x, y: mx.fmod(x, y), 15: lambda x, y: mx.pow(x, y), 16: lambda x, y: mx.maximum(x, y),
            17: lambda x, y: mx.minimum(x, y)}
    for opc, f in bins.items():
        out = np.zeros(512, np.float32)
        oracle.elementwise(f(np_tensor(a), np_tensor(b)), out)
        assert np.array_equal(out, g["functor/binary_%d" % opc]), opc


def test_oracle_black_scholes_matches_reference(oracle):
    g = GU.load()
    K, S, V, r, T = (np_tensor(g["bs/" + k]) for k in "KSVrT")
    VsqrtT = V * mx.sqrt(T)
    d1 = (mx.log(S / K) + (r + 0.5 * V * V) * T) / VsqrtT
    d2 = d1 - VsqrtT
    out = np.zeros(4096, np.float32)
    oracle.elementwise(S * mx.normcdf(d1) - K * mx.exp(-1.0 * r * T) * mx.normcdf(d2), out)
    want = g["bs/out"]
    assert np.max(np.abs(out - want)) <= 2e-5 and np.median(np.abs(out - want)) <= 1e-6


@pytest.fixture(scope="module")
def ref():
    r = H.load_ref_host()
    if r is None:
        pytest.skip("oracle/_ref/libmatx_ref_host.so not built (needs /root/reference)")
    return r


@pytest.mark.parametrize("seed", range(3))
def test_oracle_equals_live_reference_on_random_views(oracle, ref, seed):
    rng = np.random.default_rng(100 + seed)
    for dt in (np.float32, np.float64, np.int32):
        shape = tuple(int(s) for s in rng.integers(2, 12, size=int(rng.integers(1, 5))))
        x = rng.integers(-5, 6, shape).astype(dt) if dt == np.int32 else rng.random(shape).astype(dt)
        rank = len(shape)
        if dt != np.float32 and rank > 2:
            continue   # the reference wrapper instantiates ranks > 2 for fp32 only
        nd = int(rng.integers(1, rank + 1))
        dims = sorted(rng.choice(rank, nd, replace=False).tolist())
        for opn, op in [("sum", A.RED_SUM), ("max", A.RED_MAX), ("argmin", A.RED_ARGMIN), ("all", A.RED_ALL)] + (
                [] if dt == np.int32 else [("mean", A.RED_MEAN), ("var", A.RED_VAR)]):
            want, widx = ref.reduce_np(op, x, dims)
            r = getattr(mx, opn)(np_tensor(x), dims)
            out = np.zeros(r.out_shape, want.dtype)
            idx = np.zeros(r.out_shape, np.int64) if widx is not None else None
            oracle.reduce(r, out, idx)
            assert np.array_equal(out, want), (dt, shape, dims, opn)
            if widx is not None:
                assert np.array_equal(idx, widx)


def test_reference_threaded_mode_agrees_on_exact_ops(ref):
    # HostExecutor<ALL> (Thrust-OMP above 16 Ki elements) must give the same max / argmax as SINGLE
    x = np.random.default_rng(5).integers(0, 1000, 1 << 18).astype(np.float32)
    for op in (A.RED_MAX, A.RED_ARGMAX, A.RED_ANY):
        a = ref.reduce_np(op, x, [0], mode=0)
        b = ref.reduce_np(op, x, [0], mode=1)
        assert np.array_equal(a[0], b[0])
        if a[1] is not None:
            assert np.array_equal(a[1], b[1])
