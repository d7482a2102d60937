"""-m gpu parity tests of the transposing elementwise family (`ew_tr`): permuted copies `(y = x.Permute(...)).run()`
(reference bench: bench/00_operators/operators.cu:40-59; reference tests: test/00_operators/permute_test.cu,
transpose_test.cu) and expressions that mix row- and column-walking operands.  Copies are bit-exact against numpy
and against the CPU oracle; every case also asserts WHICH kernel family served it."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G

pytestmark = pytest.mark.gpu

NP = {A.F32: np.float32, A.F64: np.float64, A.C64: np.complex64, A.I32: np.int32}


def rand(rng, shape, dt):
    if dt == A.C64:
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(np.complex64)
    if dt == A.I32:
        return rng.integers(-1000, 1000, shape).astype(np.int32)
    return rng.standard_normal(shape).astype(NP[dt])


@pytest.mark.parametrize("dt", [A.F32, A.F64, A.C64, A.I32])
@pytest.mark.parametrize("shape", [(64, 64), (256, 320), (100, 37), (33, 1000), (129, 131)])
def test_transpose_2d_exact(oracle, dt, shape):
    rng = np.random.default_rng(11)
    x = rand(rng, shape, dt)
    got, want, k = G.run_elementwise(oracle, lambda t: mx.permute(t, [1, 0]), [x], shape[::-1], dt)
    assert k.startswith("ew_tr|"), k
    assert np.array_equal(got, x.T) and np.array_equal(got, want)


def test_transpose_bf16_exact(oracle):
    rng = np.random.default_rng(12)
    for shape in [(128, 128), (200, 72), (513, 257)]:
        x = rng.integers(0, 1 << 16, shape).astype(np.uint16)
        x[(x & 0x7f80) == 0x7f80] = 0x3f80  # no NaN patterns: the copy keeps bits, the oracle compares values
        got, want, k = G.run_elementwise(oracle, lambda t: mx.permute(t, [1, 0]), [x], shape[::-1], A.BF16, dtypes=[A.BF16])
        assert k.startswith("ew_tr|") and "|V8|" in k, k
        assert np.array_equal(got, x.T)


@pytest.mark.parametrize("perm,shape", [((3, 0, 2, 1), (50, 40, 6, 30)),   # the reference benchmark's permutation, scaled down
                                        ((2, 0, 1), (24, 40, 48)), ((0, 2, 1), (5, 70, 90)), ((1, 0, 2), (30, 20, 64)),
                                        ((2, 1, 0), (17, 19, 23)), ((3, 2, 1, 0), (16, 18, 20, 22))])
def test_permute_nd_exact(oracle, perm, shape):
    rng = np.random.default_rng(13)
    x = rand(rng, shape, A.F32)
    oshape = tuple(shape[p] for p in perm)
    got, want, k = G.run_elementwise(oracle, lambda t: mx.permute(t, list(perm)), [x], oshape, A.F32)
    assert np.array_equal(got, np.transpose(x, perm)) and np.array_equal(got, want)
    if perm[-1] != len(shape) - 1:   # the last dim moved: the read side walks another dim than the write side
        assert k.startswith("ew_tr|"), k
    else:                            # last dim stays: plain vectorised rows
        assert k.startswith("ew|"), k


def test_mixed_row_and_column_operands(oracle):
    rng = np.random.default_rng(14)
    a = rand(rng, (96, 200), A.F32)
    b = rand(rng, (200, 96), A.F32)
    v = rand(rng, (200,), A.F32)
    got, want, k = G.run_elementwise(oracle, lambda x, y, z: x * 2.0 + mx.permute(y, [1, 0]) - z, [a, b, v], a.shape, A.F32)
    assert k.startswith("ew_tr|"), k
    assert np.array_equal(got, want)
    assert np.allclose(got, a * np.float32(2) + b.T - v, rtol=1e-6, atol=1e-6)
    # two staged operands and an fp64 one walking rows
    c = rand(rng, (200, 96), A.F32)
    d = rng.standard_normal((96, 200))
    got, want, k = G.run_elementwise(oracle, lambda x, y, z: mx.permute(x, [1, 0]) * mx.permute(y, [1, 0]) + z, [b, c, d], (96, 200), A.F64)
    assert k.startswith("ew_tr|"), k
    assert np.array_equal(got, want)


def test_cast_on_the_way_out(oracle):
    from tests.oracle_harness import bf16_bits_to_f32
    rng = np.random.default_rng(15)
    x = rand(rng, (130, 260), A.F32)
    got, want, k = G.run_elementwise(oracle, lambda t: mx.permute(t, [1, 0]), [x], (260, 130), A.BF16)
    assert k.startswith("ew_tr|"), k
    assert np.array_equal(got, want)
    assert np.max(np.abs(bf16_bits_to_f32(got) - x.T)) <= 2 ** -8 * np.max(np.abs(x))
    got, want, k = G.run_elementwise(oracle, lambda t: mx.permute(t, [1, 0]), [x], (260, 130), A.F64)
    assert k.startswith("ew_tr|"), k
    assert np.array_equal(got, x.T.astype(np.float64))


def test_unaligned_views_take_the_scalar_paths():
    """Slices that break the 16-byte alignment of the read side, the write side, or both (odd offsets, odd pitches)."""
    import torch
    ex = G.executor()
    g = torch.Generator(device="cuda").manual_seed(16)
    big = torch.randn(301, 403, device="cuda", generator=g)
    for (r0, c0) in [(0, 0), (1, 0), (0, 1), (3, 5)]:
        x = big[r0:r0 + 256, c0:c0 + 384]                      # pitch 403 (odd), offset c0
        out_store = torch.zeros(390, 263, device="cuda")
        for (o0, o1) in [(0, 0), (1, 1), (2, 3)]:
            y = out_store[o0:o0 + 384, o1:o1 + 256]            # pitch 263 (odd)
            mx.make_tensor(y).set(mx.make_tensor(x).Permute([1, 0])).run(ex)
            ex.sync()
            assert ex.last_kernel().startswith("ew_tr|"), ex.last_kernel()
            assert torch.equal(y, x.t())
            # nothing outside the written window was touched
            assert float(out_store.abs().sum()) == pytest.approx(float(y.abs().sum()), rel=1e-6)
            out_store.zero_()


def test_permuted_left_hand_side():
    """(y.Permute({1,0}) = x): the statement is walked in the OUTPUT's layout order, so this is the same transpose."""
    import torch
    ex = G.executor()
    x = torch.randn(192, 320, device="cuda")
    y = torch.zeros(320, 192, device="cuda")
    mx.make_tensor(y).Permute([1, 0]).set(mx.make_tensor(x)).run(ex)
    ex.sync()
    assert ex.last_kernel().startswith("ew_tr|"), ex.last_kernel()
    assert torch.equal(y, x.t())


def test_small_or_thin_shapes_fall_back_to_rows(oracle):
    rng = np.random.default_rng(17)
    for shape in [(3, 1000), (1000, 3), (8, 8)]:
        x = rand(rng, shape, A.F32)
        got, want, k = G.run_elementwise(oracle, lambda t: mx.permute(t, [1, 0]), [x], shape[::-1], A.F32)
        assert k.startswith("ew|"), k
        assert np.array_equal(got, x.T)


def test_reference_bench_shape_full_size():
    """bench/00_operators/operators.cu:40-59 at full size: x {1000,200,6,300} -> y {300,1000,6,200}."""
    import torch
    ex = G.executor()
    x = torch.randn(1000, 200, 6, 300, device="cuda")
    y = torch.empty(300, 1000, 6, 200, device="cuda")
    mx.make_tensor(y).set(mx.make_tensor(x).Permute([3, 0, 2, 1])).run(ex)
    ex.sync()
    assert ex.last_kernel().startswith("ew_tr|"), ex.last_kernel()
    assert torch.equal(y, x.permute(3, 0, 2, 1))
