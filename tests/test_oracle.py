"""Pins oracle/matx_oracle.c against the known-answer vectors the reference's own tests hold for this path
(SURVEY.md section 4 / 8c).  CPU only.  Citations are to /root/reference/test/..."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests.oracle_harness import np_tensor, f32_to_bf16_bits, bf16_bits_to_f32

FLOATS = [np.float32, np.float64]


def red(oracle, r, out_dtype, idx=False, **kw):
    out = np.zeros(r.out_shape, out_dtype)
    ix = np.zeros(r.out_shape, np.int64) if idx else None
    oracle.reduce(r, out, ix, **kw)
    return (out, ix) if idx else out


@pytest.mark.parametrize("dt", FLOATS)
def test_max_min_known_answers(oracle, dt):
    # ReductionTests.cu:948-954, 973-979
    t = np.array([1, 3, 8, 2, 9, 10, 6, 7, 4, 5, 11], dt)
    assert red(oracle, mx.max(np_tensor(t)), dt) == 11
    assert red(oracle, mx.min(np_tensor(t)), dt) == 1


@pytest.mark.parametrize("dt", FLOATS)
def test_argmax_argmin_known_answers(oracle, dt):
    # ReductionTests.cu:1255-1261, 1332-1338
    t = np.array([1, 3, 8, 2, 9, 10, 6, 7, 4, 5, 11], dt)
    v, i = red(oracle, mx.argmax(np_tensor(t)), dt, idx=True)
    assert (v, i) == (11, 10)
    v, i = red(oracle, mx.argmin(np_tensor(t)), dt, idx=True)
    assert (v, i) == (1, 0)
    # non-tensor input: argmax(t + 0)  (:1263-1267)
    v, i = red(oracle, mx.argmax(np_tensor(t) + 0), dt, idx=True)
    assert (v, i) == (11, 10)
    # ReductionTests.cu:1272-1280, 1343-1351: absolute flat indices
    t2 = np.array([[2, 4, 1, 3, 5], [3, 1, 5, 2, 4]], dt)
    v, i = red(oracle, mx.argmax(np_tensor(t2), [1]), dt, idx=True)
    assert v.tolist() == [5, 5] and i.tolist() == [4, 7]
    v, i = red(oracle, mx.argmin(np_tensor(t2), [1]), dt, idx=True)
    assert v.tolist() == [1, 1] and i.tolist() == [2, 6]


def test_cub_argmax_absolute_index(oracle):
    # test/00_tensor/CUBTests.cu:328-340
    t = np.array([[1, 5, 2], [4, 3, 6]], np.float32)
    v, i = red(oracle, mx.argmax(np_tensor(t), [1]), np.float32, idx=True)
    assert v.tolist() == [5, 6] and i.tolist() == [1, 5]


def test_minmax_negative(oracle):
    # ReductionTests.cu:913-933
    t = np.array([-3, -1, -7], np.float32)
    v, i = red(oracle, mx.argmax(np_tensor(t)), np.float32, idx=True)
    assert (v, i) == (-1, 1)
    v, i = red(oracle, mx.argmin(np_tensor(t)), np.float32, idx=True)
    assert (v, i) == (-7, 2)


@pytest.mark.parametrize("sign,op", [(1, "argmax"), (-1, "argmin")])
def test_planted_extrema_6x33x33(oracle, sign, op):
    # ReductionTests.cu:1296-1311, 1367-1383
    planted = [31 * 33 + 22, 32 * 33 + 24, 19 * 33 + 12, 21 * 33 + 17, 17 * 33 + 7, 1 * 33 + 24]
    t = np.zeros((6, 33, 33), np.float32)
    for n, p in enumerate(planted):
        t[n, p // 33, p % 33] = sign
    v, i = red(oracle, getattr(mx, op)(np_tensor(t), [1, 2]), np.float32, idx=True)
    assert i.tolist() == [n * 1089 + p for n, p in enumerate(planted)]
    assert v.tolist() == [sign] * 6


def test_lowest_index_tie_break(oracle):
    # BASELINE.md section 6 probe of the reference HostExecutor: argmax -> 1, argmin -> 2
    t = np.array([3, 7, 1, 7, 1, 3, 7, 1], np.float32)
    assert red(oracle, mx.argmax(np_tensor(t)), np.float32, idx=True)[1] == 1
    assert red(oracle, mx.argmin(np_tensor(t)), np.float32, idx=True)[1] == 2


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32, np.complex64])
def test_sum_of_ones(oracle, dt):
    # ReductionTests.cu:198-311 (sum of ones = element count, every rank) and :1000-1057 (segmented)
    for shape in [(30,), (30, 40), (5, 6, 7), (3, 4, 5, 6)]:
        t = np.ones(shape, dt)
        assert red(oracle, mx.sum(np_tensor(t)), dt) == np.prod(shape)
    t = np.ones((273, 2, 4), dt)
    assert (red(oracle, mx.sum(np_tensor(t), [2]), dt) == 4).all()
    t = np.ones((3, 4, 5), dt)
    assert (red(oracle, mx.sum(np_tensor(t), [2]), dt) == 5).all()
    assert (red(oracle, mx.sum(np_tensor(t), [1, 2]), dt) == 20).all()
    rows = np.tile(np.arange(1, 6).astype(dt), (4, 1))
    assert (red(oracle, mx.sum(np_tensor(rows), [1]), dt) == 15).all()


@pytest.mark.parametrize("dt", FLOATS)
def test_mean_of_ones(oracle, dt):
    # ReductionTests.cu:1490-1561
    t = np.ones((3, 4, 5, 6), dt)
    for dims in [None, [3], [2, 3], [1, 2, 3], [0, 1], [0]]:
        r = mx.mean(np_tensor(t), dims)
        assert (red(oracle, r, dt) == 1).all()


@pytest.mark.parametrize("dt", [np.float32, np.float64, np.int32])
def test_any_all_known_answers(oracle, dt):
    # ReductionTests.cu:591-635 (any) and :688-732 (all)
    t1, t2, t3, t4 = (np.zeros(s, dt) for s in [(30,), (30, 40), (30, 40, 50), (3, 4, 5, 6)])
    t1[5] = 5
    t3[1, 1, 1] = 6
    got = [int(red(oracle, mx.any(np_tensor(t)), dt)) for t in (t4, t3, t2, t1)]
    assert got == [0, 1, 0, 1]
    a = red(oracle, mx.any(np_tensor(t3), [2]), dt)
    want = np.zeros((30, 40), dt)
    want[1, 1] = 1
    assert (a == want).all()
    o1, o2, o3, o4 = (np.ones(s, dt) for s in [(30,), (30, 40), (30, 40, 50), (3, 4, 5, 6)])
    o1[5] = 0
    o3[1, 1, 1] = 0
    got = [int(red(oracle, mx.all(np_tensor(t)), dt)) for t in (o4, o3, o2, o1)]
    assert got == [1, 0, 1, 0]
    a = red(oracle, mx.all(np_tensor(o3), [2]), dt)
    want = np.ones((30, 40), dt)
    want[1, 1] = 0
    assert (a == want).all()


@pytest.mark.parametrize("dt", FLOATS)
@pytest.mark.parametrize("ddof", [0, 1])
def test_var_std_vs_numpy(oracle, dt, ddof):
    # ReductionTests.cu:121-155 with test/test_vectors/generators/00_operators.py:98-115 (np.random.rand(100))
    x = np.random.default_rng(0).random(100).astype(dt)
    v = red(oracle, mx.var(np_tensor(x), None, ddof), dt)
    s = red(oracle, mx.stdd(np_tensor(x), None, ddof), dt)
    assert abs(v - np.var(x.astype(np.float64), ddof=ddof)) < 0.01  # the reference's own tolerance (utilities.h:66-97)
    assert abs(s - np.std(x.astype(np.float64), ddof=ddof)) < 0.01
    assert abs(v - np.var(x.astype(np.float64), ddof=ddof)) < 1e-5 * max(1.0, abs(v))


def test_var_complex_is_real(oracle):
    # ReductionTests.cu:157-187: output is the inner (real) type
    rng = np.random.default_rng(1)
    x = (rng.standard_normal(100) + 1j * rng.standard_normal(100)).astype(np.complex64)
    v = red(oracle, mx.var(np_tensor(x), None, 1), np.float32)
    assert abs(v - np.var(x.astype(np.complex128), ddof=1)) < 1e-5 * abs(v)


def test_permuted_reduce_equals_dims_form(oracle):
    # ReductionTests.cu:353-573 (PermutedReduce), reduced size
    rng = np.random.default_rng(2)
    t = rng.random((6, 5, 4, 7)).astype(np.float32)
    T = np_tensor(t)
    for name in ["sum", "mean", "max", "min", "any", "all", "prod"]:
        f = getattr(mx, name)
        a = red(oracle, f(mx.permute(T, [2, 3, 0, 1]), [2, 3]), np.float32)
        b = red(oracle, f(T, [0, 1]), np.float32)
        assert a.shape == (4, 7)
        assert np.array_equal(a, b), name
    # values of the dims form against numpy (order of summation matches a sequential loop)
    want = np.zeros((4, 7), np.float32)
    for i in range(6):
        for j in range(5):
            want = (want + t[i, j]).astype(np.float32)
    assert np.array_equal(red(oracle, mx.sum(T, [0, 1]), np.float32), want)


def test_fused_expression_sum(oracle):
    # config 1 statement: sum(a*b+c, {1}) — sequential fp32, two roundings per element (no FMA on the host path)
    rng = np.random.default_rng(3)
    a, b, c = (rng.random((8, 64)).astype(np.float32) for _ in range(3))
    got = red(oracle, mx.sum(np_tensor(a) * np_tensor(b) + np_tensor(c), [1]), np.float32)
    v = (a * b + c).astype(np.float32)
    acc = np.zeros(8, np.float32)
    for j in range(64):
        acc = (acc + v[:, j]).astype(np.float32)
    assert np.array_equal(got, acc)


def test_elementwise_black_scholes_finite(oracle):
    rng = np.random.default_rng(4)
    n = 257
    S, K = (rng.uniform(10, 100, n).astype(np.float32) for _ in range(2))
    V = rng.uniform(0.05, 0.5, n).astype(np.float32)
    r = rng.uniform(0.01, 0.1, n).astype(np.float32)
    T = rng.uniform(0.1, 2, n).astype(np.float32)
    tS, tK, tV, tr, tT = map(np_tensor, (S, K, V, r, T))
    VsqrtT = tV * mx.sqrt(tT)
    d1 = (mx.log(tS / tK) + (tr + 0.5 * tV * tV) * tT) / VsqrtT
    d2 = d1 - VsqrtT
    expr = tS * mx.normcdf(d1) - tK * mx.exp(-1.0 * tr * tT) * mx.normcdf(d2)
    out = np.zeros(n, np.float32)
    oracle.elementwise(expr, out)
    from math import erfc, exp, log, sqrt
    want = []
    for s, k, v, rr, t in zip(*(x.astype(np.float64) for x in (S, K, V, r, T))):
        vs = v * sqrt(t)
        D1 = (log(s / k) + (rr + 0.5 * v * v) * t) / vs
        D2 = D1 - vs
        N = lambda z: 0.5 * erfc(-z / sqrt(2))  # noqa: E731
        want.append(s * N(D1) - k * exp(-rr * t) * N(D2))
    assert np.allclose(out, np.array(want), rtol=2e-4, atol=2e-4)


def test_bf16_host_accumulation_stagnates(oracle):
    # SURVEY.md section 7 probe of the reference: HostExecutor sum of 1024 x bf16(0.125) returns 32 (true 128)
    bits = f32_to_bf16_bits(np.full(1024, 0.125, np.float32))
    out = np.zeros((), np.uint16)
    oracle.reduce(mx.sum(np_tensor(bits, A.BF16)), out, None, out_dtype=A.BF16, half_acc=A.BF16)
    assert bf16_bits_to_f32(out.reshape(1))[0] == 32.0
    out = np.zeros((), np.uint16)
    oracle.reduce(mx.sum(np_tensor(bits, A.BF16)), out, None, out_dtype=A.BF16)  # fp32 accumulation (the B200 path)
    assert bf16_bits_to_f32(out.reshape(1))[0] == 128.0


# ---- softmax: the reference's golden generator is scipy.special.softmax (generators/00_reductions.py:10-26) -----
@pytest.mark.parametrize("npdt,tol", [(np.float32, 2e-6), (np.float64, 1e-14)])
def test_softmax_matches_the_reference_generator(oracle, npdt, tol):
    from scipy import special
    np.random.seed(1234)                      # the generator's seed
    t1 = np.random.randn(300).astype(npdt)    # ReductionTests.cu:327-339 (size 300)
    t3 = np.random.randn(8, 30, 300).astype(npdt)
    out = np.zeros_like(t1)
    oracle.softmax(mx.softmax(np_tensor(t1)), out)
    assert np.allclose(out, special.softmax(t1.astype(np.float64)), rtol=tol, atol=0)
    out3 = np.zeros_like(t3)
    oracle.softmax(mx.softmax(np_tensor(t3), [2]), out3)
    assert np.allclose(out3, special.softmax(t3.astype(np.float64), axis=2), rtol=tol, atol=0)
    assert np.allclose(out3.sum(axis=2), 1.0, rtol=1e-5)
    # a middle axis walks both the operand and the output through the same permutation (reduce.h:404-445)
    oracle.softmax(mx.softmax(np_tensor(t3), [1]), out3)
    assert np.allclose(out3, special.softmax(t3.astype(np.float64), axis=1), rtol=tol, atol=0)
    # strided output view
    wide = np.full((8, 30, 600), -5, npdt)
    oracle.softmax(mx.softmax(np_tensor(t3), [2]), wide[:, :, ::2])
    assert np.allclose(wide[:, :, ::2], special.softmax(t3.astype(np.float64), axis=2), rtol=tol, atol=0) and (wide[:, :, 1::2] == -5).all()


def test_trace_diag_isclose_lowering(oracle):
    """trace_impl = sum(diag(in)) (transforms/reduce.h:1505-1511), diag sizes (operators/diag.h:253-268), isclose
    (operators/isclose.h: |a - b| <= atol + rtol * |b| as int) — the host-side lowering, evaluated by the oracle."""
    rng = np.random.default_rng(5)
    m = rng.standard_normal((12, 12)).astype(np.float32)
    w = np.zeros((), np.float32)
    oracle.reduce(mx.trace(np_tensor(m)), w)
    assert abs(float(w) - float(np.trace(m.astype(np.float64)))) < 1e-5
    for k in (0, 1, 5, -2):
        d = mx.diag(np_tensor(m), k)
        assert d.shape == (12 - abs(k),)
        w = np.zeros(d.shape, np.float32)
        oracle.elementwise(d * 1.0, w, A.F32)
        assert np.array_equal(w, np.diag(m, k))
    # an expression operand takes the DiagOp route: both matrix dims walk the same root dim
    w = np.zeros((), np.float32)
    oracle.reduce(mx.sum(mx.diag(np_tensor(m) * np_tensor(m))), w)
    assert abs(float(w) - float(np.sum(np.diag(m).astype(np.float64) ** 2))) < 1e-4
    stack = rng.standard_normal((3, 5, 5)).astype(np.float32)
    w = np.zeros(3, np.float32)
    oracle.reduce(mx.sum(mx.diag(np_tensor(stack)), [1]), w)
    assert np.allclose(w, np.einsum("bii->b", stack), rtol=1e-5, atol=1e-6)
    with pytest.raises(ValueError):
        mx.diag(np_tensor(np.zeros((7, 5), np.float32)), 1)   # the reference's size rule would run off the matrix
    a = rng.standard_normal((8, 9)).astype(np.float32)
    b = a + np.float32(1e-4) * (rng.random((8, 9)) > 0.5).astype(np.float32)
    w = np.zeros(a.shape, np.int32)
    oracle.elementwise(mx.isclose(np_tensor(a), np_tensor(b), 1e-5, 1e-8), w, A.I32)
    assert np.array_equal(w.astype(bool), np.abs(a - b) <= np.float32(1e-8) + np.float32(1e-5) * np.abs(b))
    flag = np.zeros((), np.int32)
    oracle.reduce(mx.all(mx.isclose(np_tensor(a), np_tensor(b), 1e-5, 1e-8)), flag, out_dtype=A.I32)
    assert int(flag) == int(np.all(w))


def test_cumsum_reference_known_answers(oracle):
    """cumsum: the reference's own known answers (test/00_tensor/CUBTests.cu:203-226 permuted int matrix, :536-575 running
    float sums per row, test/00_operators/stack_test.cu:72-80) and numpy on strided / fused inputs."""
    inv = np.array([[1, 2, 3, 4], [10, 20, 30, 40], [100, 200, 300, 400]], np.int32)
    out = np.zeros((4, 3), np.int32)
    oracle.cumsum(mx.cumsum(np_tensor(inv).Permute([1, 0])), out)
    assert np.array_equal(out, np.cumsum(inv.T, axis=1))
    a = np.array([1, 2, 3, 4, 5], np.float32)
    out = np.zeros(5, np.float32)
    oracle.cumsum(mx.cumsum(np_tensor(a)), out)
    assert out.tolist() == [1, 3, 6, 10, 15]
    outer, batches, cols = 2, 3, 40                      # CUBTests.cu:536-575: in(i,j,k) = base + (cols - k)
    x = np.zeros((outer, batches, cols), np.float32)
    for i in range(outer):
        for j in range(batches):
            x[i, j] = 1000 * i + 100 * j + (cols - np.arange(cols))
    out = np.zeros_like(x)
    oracle.cumsum(mx.cumsum(np_tensor(x)), out)
    assert np.max(np.abs(out - np.cumsum(x.astype(np.float64), axis=2))) <= 0.001
    # sequential fp32 running sum, bit for bit (std::partial_sum)
    rng = np.random.default_rng(3)
    r = rng.standard_normal((3, 257)).astype(np.float32)
    out = np.zeros_like(r)
    oracle.cumsum(mx.cumsum(np_tensor(r)), out)
    run = np.zeros(3, np.float32)
    for j in range(257):
        run = (run + r[:, j]).astype(np.float32) if j else r[:, 0].copy()
        assert np.array_equal(out[:, j], run)
    # fused operand, complex
    c = (rng.standard_normal((2, 33)) + 1j * rng.standard_normal((2, 33))).astype(np.complex64)
    out = np.zeros_like(c)
    oracle.cumsum(mx.cumsum(np_tensor(c) * 2.0), out)
    assert np.allclose(out, np.cumsum(c * 2, axis=1), rtol=1e-5, atol=1e-5)
