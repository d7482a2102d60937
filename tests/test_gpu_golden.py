"""-m gpu: the CUDA path against the committed golden vectors made by THE REFERENCE's HostExecutor
(tests/golden/reference_host.npz).  Exact ops bit-exact; sums within 1e-5 relative (fp32 / c64), 1e-12 (fp64)."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import golden_util as GU
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
NP2 = {np.dtype(np.float32): A.F32, np.dtype(np.float64): A.F64, np.dtype(np.complex64): A.C64, np.dtype(np.int32): A.I32,
       np.dtype(np.int64): A.I64}


def test_reductions_match_reference_golden(oracle):
    g = GU.load()
    n = 0
    for tag, opn, build, want, widx in GU.statements(g):
        x = g[tag + "/x"]
        got, gi, _, _, k = G.run_reduce(oracle, build, [x], NP2[want.dtype])
        if opn in GU.EXACT:
            assert np.array_equal(got, want), (tag, opn, k)
            if widx is not None:
                assert np.array_equal(gi, widx), (tag, opn, k)
        else:
            tol = 1e-12 if want.dtype == np.float64 else (1e-4 if opn == "prod" else 2e-5 if opn in ("var", "stdd") else 1e-5)
            if want.dtype == np.int32:
                assert np.array_equal(got, want), (tag, opn, k)
            else:
                assert G.rel_err(got, want) <= tol, (tag, opn, k, G.rel_err(got, want))
        n += 1
    assert n > 150


def test_fused_statements_match_reference_golden(oracle):
    g = GU.load()
    got = G.run_reduce(oracle, lambda a, b, c: mx.sum(a * b + c, [1]), [g["fma_sum/" + k] for k in "abc"], A.F32)[0]
    assert G.rel_err(got, g["fma_sum/out"]) <= 1e-5
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmax(mx.abs2(x), [1]), [g["abs2_argmax/x"]], A.F32)
    assert np.array_equal(gi, g["abs2_argmax/idx"]) and G.rel_err(got, g["abs2_argmax/val"]) <= 1e-6

    def bs(K, S, V, r, T):
        VsqrtT = V * mx.sqrt(T)
        d1 = (mx.log(S / K) + (r + 0.5 * V * V) * T) / VsqrtT
        d2 = d1 - VsqrtT
        return S * mx.normcdf(d1) - K * mx.exp(-1.0 * r * T) * mx.normcdf(d2)

    got, _, _ = G.run_elementwise(oracle, bs, [g["bs/" + k] for k in "KSVrT"], (4096,), A.F32)
    want = g["bs/out"]
    assert np.max(np.abs(got - want)) <= 2e-4
    big = want > 1.0
    assert np.max(np.abs(got[big] - want[big]) / want[big]) <= 1e-5


def test_functors_match_reference_golden(oracle):
    g = GU.load()
    a, b = g["functor/a"], g["functor/b"]
    names = {40: "neg", 41: "sqrt", 42: "rsqrt", 43: "exp", 44: "log", 45: "log2", 46: "log10", 47: "abs", 48: "abs2", 52: "sin",
             53: "cos", 54: "tan", 55: "tanh", 56: "normcdf", 60: "floor", 61: "ceil", 62: "round_", 63: "sinh", 64: "cosh",
             65: "asin", 66: "acos", 67: "atan"}
    for opc, nm in names.items():
        f = (lambda t: -t) if nm == "neg" else (lambda t, nm=nm: getattr(mx, nm)(t))
        got, _, _ = G.run_elementwise(oracle, f, [a], a.shape, A.F32)
        assert np.allclose(got, g["functor/unary_%d" % opc], rtol=2e-6, atol=1e-6), nm
    bins = {10: lambda x, y: x + y, 11: lambda x, y: x - y, 12: lambda x, y: x * y, 13: lambda x, y: x / y,
            14: lambda x, y: mx.fmod(x, y), 15: lambda x, y: mx.pow(x, y), 16: lambda x, y: mx.maximum(x, y),
            17: lambda x, y: mx.minimum(x, y)}
    for opc, f in bins.items():
        got, _, _ = G.run_elementwise(oracle, f, [a, b], a.shape, A.F32)
        assert np.allclose(got, g["functor/binary_%d" % opc], rtol=2e-6, atol=1e-6), opc
