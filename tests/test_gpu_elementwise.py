"""-m gpu parity tests of the fused elementwise executor (mxb_elementwise) against the CPU oracle.
Mirrors test/00_operators/operator_func_*_test.cu (each functor against its scalar formula), permute_test.cu,
clone_test.cu / broadcast_test.cu, and examples/black_scholes.cu of the reference."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

# fp32 device math (CUDA libdevice) vs glibc differ by a few ulp per call
RTOL, ATOL = 2e-6, 1e-6


def close(got, want, rtol=RTOL, atol=ATOL):
    ok = np.isclose(got, want, rtol=rtol, atol=atol, equal_nan=True)
    assert ok.all(), (np.asarray(got)[~ok][:5], np.asarray(want)[~ok][:5])


UNARY = ["sqrt", "rsqrt", "exp", "log", "log2", "log10", "abs", "abs2", "sin", "cos", "tan", "tanh", "sinh", "cosh", "asin",
         "acos", "atan", "normcdf", "floor", "ceil", "round_"]


@pytest.mark.parametrize("name", UNARY)
def test_unary_functors_f32(oracle, name):
    rng = np.random.default_rng(1)
    x = (rng.random(1000) * 0.98 + 0.01).astype(np.float32)  # inside every domain
    got, want, k = G.run_elementwise(oracle, lambda t: getattr(mx, name)(t), [x], x.shape, A.F32)
    close(got, want)


@pytest.mark.parametrize("name", ["sqrt", "exp", "log", "abs", "sin", "normcdf", "tanh"])
def test_unary_functors_f64(oracle, name):
    rng = np.random.default_rng(2)
    x = rng.random(777) + 0.01
    got, want, _ = G.run_elementwise(oracle, lambda t: getattr(mx, name)(t), [x], x.shape, A.F64)
    close(got, want, 1e-13, 1e-14)


def test_binary_functors(oracle):
    rng = np.random.default_rng(3)
    a = (rng.random((7, 130)) + 0.5).astype(np.float32)
    b = (rng.random((7, 130)) + 0.5).astype(np.float32)
    for f in [lambda x, y: x + y, lambda x, y: x - y, lambda x, y: x * y, lambda x, y: x / y, lambda x, y: mx.pow(x, y),
              lambda x, y: mx.fmod(x, y), lambda x, y: mx.maximum(x, y), lambda x, y: mx.minimum(x, y),
              lambda x, y: mx.atan2(x, y), lambda x, y: -x + 2.0 * y, lambda x, y: 1.0 / x - y / 3.0]:
        got, want, _ = G.run_elementwise(oracle, f, [a, b], a.shape, A.F32)
        close(got, want)
    for f in [lambda x, y: x < y, lambda x, y: x >= y, lambda x, y: x.eq(y), lambda x, y: (x > 1.0) & (y > 1.0),
              lambda x, y: (x > 1.2) | (y < 0.7), lambda x, y: ~(x > y), lambda x, y: mx.isnan(x / (y - y))]:
        got, want, _ = G.run_elementwise(oracle, f, [a, b], a.shape, A.U8)
        assert np.array_equal(got, want)


def test_integer_and_mixed_types(oracle):
    rng = np.random.default_rng(4)
    i = rng.integers(-20, 20, 500).astype(np.int32)
    j = rng.integers(1, 9, 500).astype(np.int32)
    f = rng.random(500).astype(np.float32)
    for fn, dt in [(lambda x, y, z: x + y * 3, A.I32), (lambda x, y, z: x / y, A.I32), (lambda x, y, z: x % y, A.I32),
                   (lambda x, y, z: mx.abs(x) - y, A.I32), (lambda x, y, z: x * z, A.F32),
                   (lambda x, y, z: mx.as_type(x, A.F32) / mx.as_type(y, A.F32), A.F32)]:
        got, want, _ = G.run_elementwise(oracle, fn, [i, j, f], i.shape, dt)
        if dt == A.I32:
            assert np.array_equal(got, want)
        else:
            close(got, want)


def test_complex_functors(oracle):
    rng = np.random.default_rng(5)
    x = (rng.standard_normal(600) + 1j * rng.standard_normal(600)).astype(np.complex64)
    y = (rng.standard_normal(600) + 1j * rng.standard_normal(600)).astype(np.complex64)
    r = (rng.random(600) + 0.5).astype(np.float32)
    for fn in [lambda a, b, c: a + b, lambda a, b, c: a * b, lambda a, b, c: a / b, lambda a, b, c: a * c - b / c,
               lambda a, b, c: mx.conj(a) * b, lambda a, b, c: mx.expj(c) * a]:
        got, want, _ = G.run_elementwise(oracle, fn, [x, y, r], x.shape, A.C64)
        close(got, want, 2e-5, 2e-6)
    for fn in [lambda a, b, c: mx.abs2(a), lambda a, b, c: mx.abs(a), lambda a, b, c: mx.real(a) * mx.imag(b)]:
        got, want, _ = G.run_elementwise(oracle, fn, [x, y, r], x.shape, A.F32)
        close(got, want, 2e-6, 1e-6)


def test_permute_and_broadcast(oracle):
    # permute_test.cu, clone_test.cu, broadcast_test.cu
    rng = np.random.default_rng(6)
    t = rng.random((6, 10, 12)).astype(np.float32)
    v = rng.random(12).astype(np.float32)
    got, want, _ = G.run_elementwise(oracle, lambda x, y: mx.permute(x, [2, 0, 1]) * 2.0, [t, v], (12, 6, 10), A.F32)
    assert np.array_equal(got, want) and np.array_equal(got, np.transpose(t, (2, 0, 1)) * np.float32(2))
    got, want, _ = G.run_elementwise(oracle, lambda x, y: x + y, [t, v], t.shape, A.F32)   # lower rank broadcasts
    assert np.array_equal(got, want) and np.array_equal(got, t + v)
    got, want, _ = G.run_elementwise(oracle, lambda x, y: mx.clone(y, [6, 10, mx.matxKeepDim]) - x, [t, v], t.shape, A.F32)
    assert np.array_equal(got, v - t)
    got, want, _ = G.run_elementwise(oracle, lambda x, y: mx.permute(x + y, [1, 2, 0]), [t, v], (10, 12, 6), A.F32)
    assert np.array_equal(got, np.transpose(t + v, (1, 2, 0)))


def test_scalar_assignment_and_ragged_tail(oracle):
    rng = np.random.default_rng(7)
    for n in [1, 3, 4, 5, 255, 257, 1023, 4097]:
        x = rng.random(n).astype(np.float32)
        got, want, _ = G.run_elementwise(oracle, lambda t: t * 3.0 + 1.0, [x], x.shape, A.F32)
        close(got, want)  # the device contracts t*3+1 into one FMA (as the reference's CUDA build does)
        got, want, _ = G.run_elementwise(oracle, lambda t: mx.maximum(t, 0.5) - 0.25, [x], x.shape, A.F32)
        assert np.array_equal(got, want)
    got, want, _ = G.run_elementwise(oracle, lambda t: 2.5, [np.zeros(1, np.float32)], (7, 9), A.F32)
    assert (got == 2.5).all()


def test_bf16_elementwise(oracle):
    from tests.oracle_harness import f32_to_bf16_bits, bf16_bits_to_f32
    rng = np.random.default_rng(8)
    a = f32_to_bf16_bits(rng.random(2048).astype(np.float32))
    b = f32_to_bf16_bits(rng.random(2048).astype(np.float32))
    got, want, _ = G.run_elementwise(oracle, lambda x, y: x * y + x, [a, b], a.shape, A.BF16, dtypes=[A.BF16, A.BF16])
    g, w = bf16_bits_to_f32(got), bf16_bits_to_f32(want)
    assert np.max(np.abs(g - w) / np.maximum(w, 1e-6)) <= 2 ** -7   # fused multiply-add vs two roundings, then bf16


def black_scholes(S, K, V, r, T):
    VsqrtT = V * mx.sqrt(T)
    d1 = (mx.log(S / K) + (r + 0.5 * V * V) * T) / VsqrtT
    d2 = d1 - VsqrtT
    return S * mx.normcdf(d1) - K * mx.exp(-1.0 * r * T) * mx.normcdf(d2)


def test_black_scholes_config4_shape(oracle):
    # examples/black_scholes.cu:122-138 on the finite-valued input set
    rng = np.random.default_rng(9)
    n = 1 << 16
    S, K = (rng.uniform(10, 100, n).astype(np.float32) for _ in range(2))
    V = rng.uniform(0.05, 0.5, n).astype(np.float32)
    r = rng.uniform(0.01, 0.1, n).astype(np.float32)
    T = rng.uniform(0.1, 2, n).astype(np.float32)
    got, want, k = G.run_elementwise(oracle, lambda k_, s_, v_, r_, t_: black_scholes(s_, k_, v_, r_, t_), [K, S, V, r, T], (n,), A.F32)
    assert k.startswith("ew|") and k.endswith("aot") and "|V4" in k, k
    # prices span 1e-20 .. 90.  The price is a DIFFERENCE of two terms of size ~S, so an ulp-level difference in normcdf /
    # log / exp (the engine's fp32 functions vs glibc here, vs libdevice in the reference) is worth ~1e-7 * S in the
    # price whatever the price is: the bar is 1e-6 relative to the minuend's scale S, and the north star's 1e-5
    # relative on the price itself where the subtraction does not cancel (price >= S/4).
    assert np.max(np.abs(got - want) / S) <= 1e-6
    big = want >= 0.25 * S
    assert big.sum() > 1000
    assert np.max(np.abs(got[big] - want[big]) / want[big]) <= 1e-5
