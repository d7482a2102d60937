"""-m gpu parity, round-2 additions: fp16 rows of the op x dtype matrix (fp32 accumulation, north star),
non-finite inputs (+-inf, NaN away from a row's first element) and single-element rows for sum / max / min / arg ops,
all against the CPU oracle on the same bits.  Bars: exact ops bit-exact incl. the index; sums 1e-5 (fp32 arithmetic) /
1e-2 when the OUTPUT is a 16-bit float.

NaN note: the reference's own answer for a NaN in the FIRST position of a row depends on the executor (HostExecutor:
std::max_element keeps a leading NaN, transforms/host_algorithms.h:262-296; cudaExecutor: cub::Max is order dependent,
transforms/cub.h:1120-1204), so that one case is unpinned and not tested; everywhere else a NaN never compares greater or
less and both executors skip it, which is what is checked here."""
import zlib

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import bf16_bits_to_f32, f32_to_bf16_bits
from tests.test_gpu_parity import EXACT_OPS, check

pytestmark = pytest.mark.gpu


def f16_bits(x):
    return np.ascontiguousarray(x, dtype=np.float32).astype(np.float16).view(np.uint16)


def f16_val(b):
    return b.view(np.float16).astype(np.float32)


HALF_SHAPES = [((4099,), None), ((70001,), None), ((37, 129), [1]), ((37, 129), [0]), ((5, 4096), [1]), ((300, 64), [0]),
               ((6, 33, 33), [1, 2]), ((6, 33, 33), [0]), ((16, 40, 24), [0, 2])]


@pytest.mark.parametrize("shape,dims", HALF_SHAPES)
@pytest.mark.parametrize("hdt", [A.F16, A.BF16])
def test_half_exact_ops(oracle, hdt, shape, dims):
    """max / min / argmax / argmin / any / all of fp16 and bf16 tensors: the value type is fp32 (core/half.h promotion), the
    extremum is one of the inputs, so a 16-bit output reproduces it bit for bit."""
    rng = np.random.default_rng(zlib.crc32(repr((hdt, shape, dims)).encode()))
    f = rng.integers(-6, 7, shape).astype(np.float32) * 0.25      # few distinct values: ties everywhere, exact in both formats
    f.ravel()[rng.integers(0, f.size)] = 0
    bits = f16_bits(f) if hdt == A.F16 else f32_to_bf16_bits(f).reshape(shape)
    for op in EXACT_OPS:
        for out_dt in ((hdt, A.F32) if op not in ("any", "all") else (A.F32,)):
            k = check(oracle, op, lambda t, op=op: getattr(mx, op)(t, dims), [bits], out_dt, dtypes=[hdt])
            assert "f16" in k or "bf16" in k or "E_" in k, k


@pytest.mark.parametrize("shape,dims", HALF_SHAPES)
def test_fp16_sum_family_fp32_accumulation(oracle, shape, dims):
    """sum / mean / var / stdd of fp16 inputs accumulate in fp32 (north star); fp32 output within 1e-5 of the oracle's fp32
    accumulation, fp16 output within 1e-2."""
    rng = np.random.default_rng(zlib.crc32(repr(("f16sum", shape, dims)).encode()))
    bits = f16_bits(rng.random(shape) * 0.5 + 0.25)      # mean 0.5: the 70001-element sum stays below the fp16 maximum (65504)
    for op in ("sum", "mean", "var", "stdd"):
        build = (lambda t, op=op: getattr(mx, op)(t, dims)) if op in ("sum", "mean") else (lambda t, op=op: getattr(mx, op)(t, dims, 1))
        n_red = np.prod([shape[d] for d in (dims if dims is not None else range(len(shape)))])
        if op in ("var", "stdd") and n_red < 2:
            continue
        got, _, want, _, k = G.run_reduce(oracle, build, [bits], A.F32, dtypes=[A.F16])
        assert G.rel_err(got, want) <= 1e-5, (op, k, G.rel_err(got, want))
        got, _, want, _, k = G.run_reduce(oracle, build, [bits], A.F16, dtypes=[A.F16])
        assert G.rel_err(f16_val(got), f16_val(want)) <= 1e-2, (op, k)
    # against fp64 truth too: fp32 accumulation does not stagnate where a running fp16 sum would (1024 x 0.125 -> 32 there)
    if dims is None:
        got, _, _, _, _ = G.run_reduce(oracle, lambda t: mx.sum(t), [bits], A.F32, dtypes=[A.F16])
        truth = f16_val(bits).astype(np.float64).sum()
        assert abs(float(got) - truth) <= 1e-5 * truth


def test_fp16_fused_elementwise_and_reduce(oracle):
    rng = np.random.default_rng(16)
    a, b, c = (f16_bits(rng.random((64, 2048))) for _ in range(3))
    got, want, k = G.run_elementwise(oracle, lambda x, y, z: x * y + z, [a, b, c], (64, 2048), A.F16, dtypes=[A.F16] * 3)
    assert np.array_equal(got, want), k          # fp32 arithmetic, one rounding to fp16 at the store: bit-exact
    got, want, k = G.run_elementwise(oracle, lambda x, y, z: x * y + z, [a, b, c], (64, 2048), A.F32, dtypes=[A.F16] * 3)
    assert np.array_equal(got, want), k
    got, _, want, _, k = G.run_reduce(oracle, lambda x, y, z: mx.sum(x * y + z, [1]), [a, b, c], A.F32, dtypes=[A.F16] * 3)
    assert G.rel_err(got, want) <= 1e-5, k
    # bf16 twin of the same statement
    a, b, c = (f32_to_bf16_bits(rng.random((64, 2048))).reshape(64, 2048) for _ in range(3))
    got, want, k = G.run_elementwise(oracle, lambda x, y, z: x * y + z, [a, b, c], (64, 2048), A.BF16, dtypes=[A.BF16] * 3)
    assert np.array_equal(got, want), k
    assert bf16_bits_to_f32(got).shape == (64, 2048)


# ---- non-finite values and degenerate rows ---------------------------------------------------------------------
def _nonfinite_rows(rng, rows, cols):
    x = rng.standard_normal((rows, cols)).astype(np.float32)
    inf = np.float32(np.inf)
    x[0, cols // 3] = inf                      # +inf: max, argmax; the sum becomes +inf
    x[1, cols // 2] = -inf                     # -inf: min, argmin; the sum becomes -inf
    x[2, 1] = inf
    x[2, cols - 1] = inf                       # two +inf: lowest index wins
    x[3, :] = -inf                             # nothing compares greater than the identity: first element (std::max_element)
    x[4, :] = inf
    x[5, 2] = inf
    x[5, cols - 2] = -inf                      # inf + -inf = NaN in the sum
    x[6, cols // 2] = np.nan                   # NaN away from the first position: skipped by max / min / arg ops
    x[7, 1:] = np.nan                          # every element after the first is NaN: the first element is the answer
    return x


@pytest.mark.parametrize("cols", [1, 7, 129, 4096, 70001])
def test_nonfinite_rows_match_oracle(oracle, cols):
    rng = np.random.default_rng(cols)
    x = _nonfinite_rows(rng, 12, max(cols, 8)) if cols >= 7 else rng.standard_normal((12, cols)).astype(np.float32)
    if cols < 7:
        x[3, :] = -np.inf
        x[4, :] = np.inf
    for op in ("max", "min", "argmax", "argmin"):
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t, [1]), [x], A.F32)
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.sum(t, [1]), [x], A.F32)
    assert np.array_equal(np.isnan(got), np.isnan(want)) and np.array_equal(np.isinf(got), np.isinf(want)), k
    fin = np.isfinite(want)
    assert np.array_equal(np.sign(got[~fin & ~np.isnan(want)]), np.sign(want[~fin & ~np.isnan(want)]))
    assert np.allclose(got[fin], want[fin], rtol=1e-4, atol=1e-4)     # zero-mean rows: an absolute bar for the finite ones
    # full-tensor forms of the same data (grid-level combine carries the non-finite values too)
    flat = x[[0, 1, 2, 5, 6, 8, 9, 10, 11]].ravel().copy()
    for op in ("max", "min", "argmax", "argmin"):
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t), [flat], A.F32)


@pytest.mark.parametrize("dt", [A.F32, A.F64, A.I32, A.C64])
def test_single_element_reductions(oracle, dt):
    """Rows of one element and one-element tensors: every reduction returns that element (index 0 / b), var with
    ddof = 0 returns 0 (ReductionTests.cu:121-187 runs ddof 0 and 1)."""
    rng = np.random.default_rng(dt)
    npdt = {A.F32: np.float32, A.F64: np.float64, A.I32: np.int32, A.C64: np.complex64}[dt]
    col = (rng.integers(-9, 9, (257, 1)) + (2j if dt == A.C64 else 0)).astype(npdt)
    one = col[:1, 0].copy()
    ops = ["sum", "any", "all"] + ([] if dt == A.C64 else ["max", "min", "argmax", "argmin"]) + (["prod"] if dt in (A.F32, A.C64) else [])
    for op in ops:
        exact = op in EXACT_OPS
        for arr, dims in ((col, [1]), (one, None)):
            got, gi, want, wi, k = G.run_reduce(oracle, lambda t, op=op: getattr(mx, op)(t, dims), [arr], dt)
            assert np.array_equal(got, want), (op, k)
            if gi is not None:
                assert np.array_equal(gi, wi) and (gi.ravel() == np.arange(gi.size)).all(), (op, k)
            assert exact or op in ("sum", "prod")
    if dt in (A.F32, A.C64, A.F64):
        odt = A.F32 if dt == A.C64 else dt
        got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.var(t, [1], 0), [col], odt)
        assert np.array_equal(got, want) and not got.any(), k
        got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.mean(t, [1]), [col], dt)
        assert np.array_equal(got, want), k


def test_extreme_magnitudes_and_signed_zero(oracle):
    """largest / smallest finite fp32, denormals, -0.0 vs +0.0 (a tie by value: the lowest index wins whichever sign)."""
    x = np.array([0.0, -0.0, np.finfo(np.float32).tiny / 4, -np.finfo(np.float32).tiny / 4, np.finfo(np.float32).max,
                  -np.finfo(np.float32).max, 1.0, -1.0] * 33, np.float32)
    for op in ("max", "min", "argmax", "argmin"):
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t), [x], A.F32)
    z = np.array([-0.0, 0.0, -0.0, 0.0] * 100, np.float32)
    for op in ("argmax", "argmin"):
        got, gi, want, wi, _ = G.run_reduce(oracle, lambda t, op=op: getattr(mx, op)(t), [z], A.F32)
        assert gi == wi == 0
    for op in ("max", "min"):
        got, _, want, _, _ = G.run_reduce(oracle, lambda t, op=op: getattr(mx, op)(t), [z], A.F32)
        assert got == want == 0.0


# ---- argminmax: both extrema from one read --------------------------------------------------------------------------
ARGMM_CASES = [((4099,), None, np.float32), ((1 << 22,), None, np.float32), ((37, 129), [1], np.float32), ((5, 70001), [1], np.float32),
               ((300, 64), [0], np.float32), ((6, 33, 33), [1, 2], np.float32), ((64, 4096), [1], np.float64),
               ((9, 5000), [1], np.int32), ((2048, 40), [1], np.float32)]


@pytest.mark.parametrize("shape,dims,npdt", ARGMM_CASES)
def test_argminmax_equals_argmin_and_argmax_of_the_oracle(oracle, shape, dims, npdt):
    """(mtie(mn, imn, mx, imx) = argminmax(x, dims)): values and absolute flat indices bit-exact against the oracle's argmin
    and argmax of the same bits (operators/argminmax.h; ReductionTests.cu ArgMinMax), ties -> lowest index on both sides."""
    import torch
    rng = np.random.default_rng(zlib.crc32(repr(("amm", shape, dims)).encode()))
    x = rng.integers(-50, 50, shape).astype(npdt)             # few distinct values: the extrema repeat
    dt = {np.float32: A.F32, np.float64: A.F64, np.int32: A.I32}[npdt]
    _, _, wmin, wimin, _ = G.run_reduce(oracle, lambda t: mx.argmin(t, dims), [x], dt)
    _, _, wmax, wimax, _ = G.run_reduce(oracle, lambda t: mx.argmax(t, dims), [x], dt)
    dx = torch.from_numpy(x).cuda()
    oshape = wmin.shape
    tdt = {np.float32: torch.float32, np.float64: torch.float64, np.int32: torch.int32}[npdt]
    mn, mxv = torch.zeros(oshape, dtype=tdt, device="cuda"), torch.zeros(oshape, dtype=tdt, device="cuda")
    imn, imx = torch.full(oshape, -1, dtype=torch.int64, device="cuda"), torch.full(oshape, -1, dtype=torch.int64, device="cuda")
    ex = mx.CudaExecutor()
    mx.mtie(*(mx.make_tensor(t) for t in (mn, imn, mxv, imx))).set(mx.argminmax(mx.make_tensor(dx), dims)).run(ex)
    ex.sync()
    k = ex.last_kernel()
    assert np.array_equal(mn.cpu().numpy(), wmin) and np.array_equal(mxv.cpu().numpy(), wmax), k
    assert np.array_equal(imn.cpu().numpy(), wimin) and np.array_equal(imx.cpu().numpy(), wimax), k
    if dims != [0]:
        assert "argminmax" in k, k        # the fused instance served it (a strided reduce dim runs argmin + argmax)


def test_argminmax_of_a_fused_expression_and_mismatched_outputs(oracle):
    import torch
    rng = np.random.default_rng(77)
    a, b = rng.standard_normal((128, 3000)).astype(np.float32), rng.standard_normal((128, 3000)).astype(np.float32)
    _, _, wmin, wimin, _ = G.run_reduce(oracle, lambda s, t: mx.argmin(mx.abs(s) - t, [1]), [a, b], A.F32)
    _, _, wmax, wimax, _ = G.run_reduce(oracle, lambda s, t: mx.argmax(mx.abs(s) - t, [1]), [a, b], A.F32)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    ta, tb = mx.make_tensor(da), mx.make_tensor(db)
    ex = mx.CudaExecutor()
    mn, mxv = torch.zeros(128, device="cuda"), torch.zeros(128, device="cuda")
    imn, imx = torch.zeros(128, dtype=torch.int64, device="cuda"), torch.zeros(128, dtype=torch.int64, device="cuda")
    mx.mtie(*(mx.make_tensor(t) for t in (mn, imn, mxv, imx))).set(mx.argminmax(mx.abs(ta) - tb, [1])).run(ex)
    ex.sync()
    assert "argminmax" in ex.last_kernel()
    assert np.array_equal(mn.cpu().numpy(), wmin) and np.array_equal(imn.cpu().numpy(), wimin)
    assert np.array_equal(mxv.cpu().numpy(), wmax) and np.array_equal(imx.cpu().numpy(), wimax)
    # max outputs strided differently from the min outputs: the library runs argmin + argmax (same answers)
    wide = torch.zeros(128, 2, device="cuda")
    mx2 = wide[:, 0]
    mx.mtie(mx.make_tensor(mn), mx.make_tensor(imn), mx.make_tensor(mx2), mx.make_tensor(imx)).set(mx.argminmax(mx.abs(ta) - tb, [1])).run(ex)
    ex.sync()
    assert "argminmax" not in ex.last_kernel()
    assert np.array_equal(mx2.cpu().numpy(), wmax) and np.array_equal(mn.cpu().numpy(), wmin)
