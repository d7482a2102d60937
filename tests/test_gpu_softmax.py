"""-m gpu parity tests for softmax: mxb_softmax through the C ABI against the CPU oracle's
restatement of softmax_impl (transforms/reduce.h:362-445) and against scipy.special.softmax, which is the reference's
own golden generator (test/test_vectors/generators/00_reductions.py:10-26; ReductionTests.cu:319-352 compares to
0.01 — the bar here is 1e-5 relative for fp32, 1e-12 for fp64, 2^-8 for bf16)."""
import numpy as np
import pytest
from scipy import special

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import np_tensor, f32_to_bf16_bits, bf16_bits_to_f32

pytestmark = pytest.mark.gpu


def run_softmax(oracle, build, arrays, out_dtype=A.F32, dtypes=None, out_view=None):
    """build(*tensors) -> SoftmaxExpr.  Returns (got, want, kernel name, launches)."""
    import torch
    dtypes = dtypes or [None] * len(arrays)
    dev = [G.to_dev(a, d) for a, d in zip(arrays, dtypes)]
    r = build(*[mx.make_tensor(t) for t in dev])
    tdt = {A.BF16: torch.bfloat16, A.F16: torch.float16, A.F64: torch.float64}.get(out_dtype, torch.float32)
    out_d = torch.full(r.out_shape, -77, dtype=tdt, device="cuda")
    view = out_view(out_d) if out_view else out_d
    ex = G.executor()
    n0 = ex.launch_count()
    mx.make_tensor(view).set(r).run(ex)
    ex.sync()
    k, nl = ex.last_kernel(), ex.launch_count() - n0
    got = G.from_dev(out_d, out_dtype)
    r2 = build(*[np_tensor(np.ascontiguousarray(a), d) for a, d in zip(arrays, dtypes)])
    if out_dtype in (A.BF16, A.F16):
        want = f32_to_bf16_bits(np.full(r2.out_shape, -77, np.float32))
    else:
        want = np.full(r2.out_shape, -77, G._NP_OF[out_dtype])
    oracle.softmax(r2, out_view(want) if out_view else want, out_dtype=out_dtype)
    return got, want, k, nl


def close(got, want, tol):
    return np.allclose(got, want, rtol=tol, atol=tol * 1e-3, equal_nan=True)


@pytest.mark.parametrize("shape,dims,family,launches", [
    ((1, 1085, 8, 16), [3], "softmax_group", 1),       # the reference bench's warm-up statement (reduction.cu:18)
    ((300,), None, "softmax_group", 1),                # ReductionTests.cu:336
    ((8, 30, 300), [2], "softmax_group", 1),           # ReductionTests.cu:345
    ((37, 512), [1], "softmax_group", 1),
    ((37, 1024), [1], "softmax_reg", 1),
    ((64, 4096), [1], "softmax_reg", 1),
    ((9, 16384), [1], "softmax_reg", 1),
    ((33, 301), [1], "softmax_reg", 1),                # ragged: scalar instance
    ((5, 7, 11), [1, 2], "softmax_group", 1),          # two trailing dims collapse into one run
    ((6, 100000), [1], "ew", 2),                       # row beyond the register budget: statistics + apply
    ((1, 1085, 8, 16), None, "ew", 2),                 # whole tensor (reduction.cu:22)
    ((40, 96, 64), [1], "ew", 2),                      # middle axis: strided reduce dim rides reduce_outer
    ((300, 16, 8), [0], "ew", 2),
])
def test_softmax_shapes(oracle, shape, dims, family, launches):
    rng = np.random.default_rng(17)
    x = (rng.standard_normal(shape) * 3).astype(np.float32)
    got, want, k, nl = run_softmax(oracle, lambda t: mx.softmax(t, dims), [x])
    assert family in k and nl == launches, (k, nl)
    # the bar: 1e-5 relative against fp64 truth (scipy is the reference's own golden generator)
    axis = None if dims is None else tuple(dims)
    assert np.allclose(got, special.softmax(x.astype(np.float64), axis=axis), rtol=1e-5, atol=1e-9)
    # the oracle restates the reference's SEQUENTIAL fp32 sum, which is itself off by up to ~R * 2^-25 (measured 1.8e-4
    # at R = 138880): agreement with it is checked to that bound, not tighter
    R = int(np.prod([shape[d] for d in (dims if dims is not None else range(len(shape)))]))
    assert close(got, want, max(1e-5, R * 2.0 ** -25)), (k, np.abs(got - want).max())


def test_softmax_f64_and_bf16(oracle):
    rng = np.random.default_rng(18)
    x = rng.standard_normal((50, 777)) * 5
    got, want, k, _ = run_softmax(oracle, lambda t: mx.softmax(t, [1]), [x], out_dtype=A.F64)
    assert "softmax" in k and close(got, want, 1e-12), k
    assert np.allclose(got, special.softmax(x, axis=1), rtol=1e-12, atol=0)
    xb = f32_to_bf16_bits(rng.standard_normal((64, 2048)).astype(np.float32))
    got, want, k, _ = run_softmax(oracle, lambda t: mx.softmax(t, [1]), [xb], out_dtype=A.BF16, dtypes=[A.BF16])
    assert "softmax_reg" in k and "|V8|" in k, k
    g, w = bf16_bits_to_f32(got), bf16_bits_to_f32(want)
    truth = special.softmax(bf16_bits_to_f32(xb).astype(np.float64), axis=1)
    assert np.abs(g - truth).max() <= 2.0 ** -8 * truth.max() and np.allclose(g, w, rtol=2.0 ** -7, atol=1e-6)


def test_softmax_of_a_fused_expression_and_strided_views(oracle):
    rng = np.random.default_rng(19)
    a, b, c = (rng.standard_normal((128, 512)).astype(np.float32) for _ in range(3))
    got, want, k, nl = run_softmax(oracle, lambda x, y, z: mx.softmax(x * y + z, [1]), [a, b, c])
    assert "softmax_group" in k and nl == 1 and close(got, want, 1e-5), k
    # input: every other row of a wider matrix; output: a column window of a wider buffer
    big = rng.standard_normal((64, 1000)).astype(np.float32)
    got, want, k, nl = run_softmax(oracle, lambda t: mx.softmax(mx.Tensor(t.data_ptr, t.dtype, (32, 400), (2000, 1), keepalive=t), [1]), [big])
    assert nl == 1 and close(got, want, 1e-5), k
    got, want, k, nl = run_softmax(oracle, lambda t: mx.softmax(mx.permute(t, [1, 0]), [1]), [big])   # columns of `big`
    assert nl == 2 and close(got, want, 1e-5), k


def test_softmax_infinities(oracle):
    x = np.random.default_rng(20).standard_normal((4, 64)).astype(np.float32)
    x[0, 3] = -np.inf
    x[1, :] = -np.inf          # exp(-inf - -inf): NaN row, as the reference's arithmetic gives
    x[2, 0] = -np.inf
    x[3, 5] = 80.0
    got, want, k, _ = run_softmax(oracle, lambda t: mx.softmax(t, [1]), [x])
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.isnan(got[1]).all()
    assert close(got, want, 1e-5)
    long = np.random.default_rng(21).standard_normal((3, 70000)).astype(np.float32)
    long[0, 0] = -np.inf
    long[1, ::2] = -np.inf
    long[2, 69999] = 60.0
    got, want, k, nl = run_softmax(oracle, lambda t: mx.softmax(t, [1]), [long])
    assert nl == 2 and close(got, want, 1e-4) and not np.isnan(got).any(), k


def test_softmax_errors():
    import torch
    ex = G.executor()
    x = torch.zeros((4, 8), dtype=torch.complex64, device="cuda")
    o = torch.zeros((4, 8), dtype=torch.complex64, device="cuda")
    with pytest.raises(A.MatxB200Error) as ei:
        mx.make_tensor(o).set(mx.softmax(mx.make_tensor(x), [1])).run(ex)
    assert ei.value.status == A.ERR_NOT_SUPPORTED
    xf = torch.zeros((4, 8), device="cuda")
    with pytest.raises(A.MatxB200Error) as ei:
        mx.make_tensor(torch.zeros((4, 9), device="cuda")).set(mx.softmax(mx.make_tensor(xf), [1])).run(ex)
    assert ei.value.status == A.ERR_SIZE
