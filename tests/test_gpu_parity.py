"""-m gpu parity tests: the CUDA path, called through the C ABI, against the CPU oracle on the same seeded
inputs, at sizes the oracle finishes in seconds.  Bars (north star): min / max / argmin / argmax / any / all
bit-exact with lowest-index ties; sum / mean / var within 1e-5 relative for fp32 (fp64: 1e-12), 1e-2 for bf16.
The test bodies mirror test/00_operators/ReductionTests.cu of the reference (cited per test)."""
import zlib

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import f32_to_bf16_bits, bf16_bits_to_f32

pytestmark = pytest.mark.gpu

TOL = {A.F32: 1e-5, A.F64: 1e-12, A.C64: 1e-5, A.BF16: 1e-2}
NPDT = {A.F32: np.float32, A.F64: np.float64, A.C64: np.complex64, A.I32: np.int32, A.I64: np.int64}


def data(rng, shape, dt, ties=False):
    if dt == A.C64:
        # non-zero mean: a relative bound on a sum is meaningless under cancellation
        return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape) + (4 + 3j)).astype(np.complex64)
    if dt in (A.I32, A.I64):
        return rng.integers(-50, 50, shape).astype(NPDT[dt])
    if ties:
        return rng.integers(0, 7, shape).astype(NPDT[dt])  # few distinct values: ties everywhere
    return rng.random(shape).astype(NPDT[dt])


EXACT_OPS = ["max", "min", "argmax", "argmin", "any", "all"]
SUM_OPS = ["sum", "mean", "prod", "var", "stdd"]


def check(oracle, opname, build, arrays, out_dt, dtypes=None, tol=None, expect_kernel=None):
    got, gi, want, wi, k = G.run_reduce(oracle, build, arrays, out_dt, dtypes)
    if expect_kernel:
        assert expect_kernel in k, k
    if opname in EXACT_OPS:
        assert np.array_equal(got, want), (opname, k)
        if gi is not None:
            assert np.array_equal(gi, wi), (opname, k, gi.ravel()[:8], wi.ravel()[:8])
    else:
        if out_dt == A.BF16:
            got, want = bf16_bits_to_f32(got), bf16_bits_to_f32(want)
        err = G.rel_err(got, want)
        assert err <= (tol or TOL[out_dt]), (opname, k, err)
    return k


# ---- known-answer vectors of the reference tests, on the device ------------------------------------------
def test_known_answers_device(oracle):
    # ReductionTests.cu:948-954, 973-979, 1255-1261, 1272-1280, 1332-1351; CUBTests.cu:328-340
    t = np.array([1, 3, 8, 2, 9, 10, 6, 7, 4, 5, 11], np.float32)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmax(x), [t], A.F32)
    assert (got, gi) == (11, 10)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmin(x), [t], A.F32)
    assert (got, gi) == (1, 0)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmax(x + 0), [t], A.F32)
    assert (got, gi) == (11, 10)
    t2 = np.array([[2, 4, 1, 3, 5], [3, 1, 5, 2, 4]], np.float32)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmax(x, [1]), [t2], A.F32)
    assert got.tolist() == [5, 5] and gi.tolist() == [4, 7]
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmin(x, [1]), [t2], A.F32)
    assert got.tolist() == [1, 1] and gi.tolist() == [2, 6]
    t3 = np.array([[1, 5, 2], [4, 3, 6]], np.float32)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmax(x, [1]), [t3], A.F32)
    assert got.tolist() == [5, 6] and gi.tolist() == [1, 5]
    t4 = np.array([-3, -1, -7], np.float32)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: mx.argmax(x), [t4], A.F32)
    assert (got, gi) == (-1, 1)
    tie = np.array([3, 7, 1, 7, 1, 3, 7, 1], np.float32)
    assert G.run_reduce(oracle, lambda x: mx.argmax(x), [tie], A.F32)[1] == 1
    assert G.run_reduce(oracle, lambda x: mx.argmin(x), [tie], A.F32)[1] == 2


@pytest.mark.parametrize("sign,op", [(1, "argmax"), (-1, "argmin")])
def test_planted_extrema_device(oracle, sign, op):
    # ReductionTests.cu:1296-1311, 1367-1383
    planted = [31 * 33 + 22, 32 * 33 + 24, 19 * 33 + 12, 21 * 33 + 17, 17 * 33 + 7, 1 * 33 + 24]
    t = np.zeros((6, 33, 33), np.float32)
    for n, p in enumerate(planted):
        t[n, p // 33, p % 33] = sign
    got, gi, _, _, _ = G.run_reduce(oracle, lambda x: getattr(mx, op)(x, [1, 2]), [t], A.F32)
    assert gi.tolist() == [n * 1089 + p for n, p in enumerate(planted)]
    assert got.tolist() == [sign] * 6


def test_any_all_device(oracle):
    # ReductionTests.cu:591-635, 688-732
    for dt, a in [(np.float32, A.F32), (np.int32, A.I32)]:
        t1, t3, t4 = np.zeros(30, dt), np.zeros((30, 40, 50), dt), np.zeros((3, 4, 5, 6), dt)
        t1[5] = 5
        t3[1, 1, 1] = 6
        assert G.run_reduce(oracle, lambda x: mx.any(x), [t4], a)[0] == 0
        assert G.run_reduce(oracle, lambda x: mx.any(x), [t3], a)[0] == 1
        assert G.run_reduce(oracle, lambda x: mx.any(x), [t1], a)[0] == 1
        got = G.run_reduce(oracle, lambda x: mx.any(x, [2]), [t3], a)[0]
        want = np.zeros((30, 40), dt)
        want[1, 1] = 1
        assert np.array_equal(got, want)
        o3 = np.ones((30, 40, 50), dt)
        o3[1, 1, 1] = 0
        assert G.run_reduce(oracle, lambda x: mx.all(x), [o3], a)[0] == 0
        got = G.run_reduce(oracle, lambda x: mx.all(x, [2]), [o3], a)[0]
        assert np.array_equal(got, 1 - want)


# ---- op x dtype x layout matrix ----------------------------------------------------------------------------
SHAPES = [
    ((1,), None), ((33,), None), ((1000,), None), ((4099,), None), ((70001,), None),
    ((37, 129), [1]), ((37, 129), [0]), ((5, 4096), [1]), ((3, 9000), [1]), ((300, 64), [1]), ((300, 64), [0]),
    ((6, 33, 33), [1, 2]), ((6, 33, 33), [2]), ((6, 33, 33), [0]), ((6, 33, 33), [0, 2]),
    ((5, 6, 7, 8), [2, 3]), ((5, 6, 7, 8), [0, 1]), ((5, 6, 7, 8), [1, 3]), ((5, 6, 7, 8), None),
]


@pytest.mark.parametrize("shape,dims", SHAPES)
@pytest.mark.parametrize("dt", [A.F32, A.F64, A.I32])
def test_exact_ops_matrix(oracle, dt, shape, dims):
    rng = np.random.default_rng(zlib.crc32(repr((dt, shape, dims)).encode()))
    x = data(rng, shape, dt, ties=True)
    x.ravel()[rng.integers(0, x.size)] = 0  # make all() interesting
    for op in EXACT_OPS:
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t, dims), [x], dt)


@pytest.mark.parametrize("shape,dims", SHAPES)
@pytest.mark.parametrize("dt", [A.F32, A.F64, A.C64])
def test_sum_ops_matrix(oracle, dt, shape, dims):
    rng = np.random.default_rng(zlib.crc32(repr((dt, shape, dims, 1)).encode()))
    x = data(rng, shape, dt)
    if dt != A.C64:
        x = x + NPDT[dt](0.5)
    n_red = int(np.prod([shape[d] for d in (dims if dims is not None else range(len(shape)))]))
    for op in SUM_OPS:
        if op == "prod":
            if dt == A.C64:   # unit-modulus factors with a little gain: the product stays finite
                y = (np.exp(1j * rng.random(shape) * 6.28) * (1 + 0.01 * rng.standard_normal(shape))).astype(np.complex64)
            else:
                y = NPDT[dt](0.9) + x * NPDT[dt](0.2) - NPDT[dt](0.1)
            if n_red > 300:
                continue  # overflows / underflows in any order
            check(oracle, op, lambda t: mx.prod(t, dims), [y], dt, tol=1e-4 if dt != A.F64 else 1e-10)
            continue
        if op in ("var", "stdd"):
            if n_red < 2:
                continue
            out_dt = A.F32 if dt == A.C64 else dt
            check(oracle, op, lambda t, op=op: getattr(mx, op)(t, dims, 1), [x], out_dt, tol=2e-5 if dt != A.F64 else 1e-10)
            check(oracle, op, lambda t, op=op: getattr(mx, op)(t, dims, 0), [x], out_dt, tol=2e-5 if dt != A.F64 else 1e-10)
            continue
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t, dims), [x], dt)


def test_permuted_reduce(oracle):
    # ReductionTests.cu:353-573: every reduction of permute(t4,{2,3,0,1}) over {2,3} equals the {0,1} form
    rng = np.random.default_rng(7)
    t = (rng.random((12, 10, 9, 16)) + 0.5).astype(np.float32)
    for op in ["sum", "mean", "max", "min", "any", "all", "var", "stdd"]:
        f = getattr(mx, op)
        a = G.run_reduce(oracle, lambda x: f(mx.permute(x, [2, 3, 0, 1]), [2, 3]), [t], A.F32)
        b = G.run_reduce(oracle, lambda x: f(x, [0, 1]), [t], A.F32)
        assert np.array_equal(a[0], b[0]), op
        assert G.rel_err(a[0], a[2]) <= 2e-5, op
    for op in ["argmax", "argmin"]:
        f = getattr(mx, op)
        a = G.run_reduce(oracle, lambda x: f(mx.permute(x, [2, 3, 0, 1]), [2, 3]), [t], A.F32)
        assert np.array_equal(a[0], a[2]) and np.array_equal(a[1], a[3]), op


def test_kernel_families_are_selected(oracle):
    rng = np.random.default_rng(8)
    x = data(rng, (64, 4096), A.F32)
    k = check(oracle, "sum", lambda t: mx.sum(t, [1]), [x], A.F32)
    assert k.startswith("red_inner") and "|T0" in k and "|V8|U4" in k and k.endswith("aot"), k   # few 16 KB rows: a CTA per row, 32-byte loads for sums
    k = check(oracle, "argmax", lambda t: mx.argmax(t, [1]), [x], A.F32)
    assert k.startswith("red_inner") and "|V4|U4" in k and k.endswith("aot"), k   # (value, index) ops keep 16-byte loads
    x = data(rng, (1500, 4096), A.F32)
    k = check(oracle, "sum", lambda t: mx.sum(t, [1]), [x], A.F32)
    assert k.startswith("red_inner") and "|T1" in k, k                                          # many 16 KB rows: a warp per row
    x = data(rng, (16, 16384), A.F32)
    k = check(oracle, "sum", lambda t: mx.sum(t, [1]), [x], A.F32)
    assert k.startswith("red_inner") and "|T0" in k, k                                          # 64 KB rows: a CTA per row
    x = data(rng, (512, 256), A.F32)
    k = check(oracle, "sum", lambda t: mx.sum(t, [1]), [x], A.F32)
    assert k.startswith("red_inner") and "|T1" in k, k
    k = check(oracle, "sum", lambda t: mx.sum(t, [0]), [x], A.F32)
    assert k.startswith("red_outer") and "|V4" in k, k
    x = data(rng, (3, 1 << 20), A.F32)  # few long rows: several CTAs per row + in-launch grid combine
    k = check(oracle, "argmax", lambda t: mx.argmax(t, [1]), [x.round(2)], A.F32)
    assert k.startswith("red_inner") and "|T0" in k, k
    x = data(rng, (257, 129), A.F32)    # odd row pitch: scalar path
    k = check(oracle, "sum", lambda t: mx.sum(t, [1]), [x], A.F32)
    assert "|V1" in k, k


def test_grid_combine_is_deterministic(oracle):
    import torch
    rng = np.random.default_rng(9)
    x = G.to_dev(rng.random(1 << 22).astype(np.float32))
    ex = G.executor()
    outs = []
    for _ in range(5):
        o = torch.zeros((), dtype=torch.float32, device="cuda")
        mx.make_tensor(o).set(mx.sum(mx.make_tensor(x))).run(ex)
        ex.sync()
        outs.append(o.item())
    assert len(set(outs)) == 1
    truth = float(x.double().sum().item())
    assert abs(outs[0] - truth) <= 1e-5 * truth


def test_sliced_unaligned_views(oracle):
    rng = np.random.default_rng(10)
    x = data(rng, (40, 1031), A.F32, ties=True)
    for op in ["sum", "max", "argmax", "argmin", "all"]:
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t.Slice([3, 1], [37, 1030]), [1]), [x], A.F32)
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t.Slice([0, 5], [40, 900])), [x], A.F32)


def test_fused_fma_sum_config1_shape(oracle):
    # config 1 at reduced height: (out = sum(a*b+c, {1})).run(exec)
    rng = np.random.default_rng(11)
    a, b = (rng.random((96, 4096)).astype(np.float32) for _ in range(2))
    c = (rng.random((96, 4096)) - 0.5).astype(np.float32)
    k = check(oracle, "sum", lambda A_, B_, C_: mx.sum(A_ * B_ + C_, [1]), [a, b, c], A.F32)
    assert k.endswith("aot"), k
    # max / argmax of the fused expression: the device contracts a*b+c into one FMA (so does the reference's CUDA
    # build), the host oracle rounds twice; compare with the FMA value computed exactly in fp64
    fma = (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(np.float32)
    got, gi, _, _, _ = G.run_reduce(oracle, lambda A_, B_, C_: mx.argmax(A_ * B_ + C_, [1]), [a, b, c], A.F32)
    assert np.array_equal(got, fma.max(1)) and np.array_equal(gi, fma.argmax(1) + np.arange(96) * 4096)
    got = G.run_reduce(oracle, lambda A_, B_, C_: mx.max(A_ * B_ + C_, [1]), [a, b, c], A.F32)[0]
    assert np.array_equal(got, fma.max(1))


def test_complex_rows_config3_shape(oracle):
    # config 3 at reduced height: mean / var(ddof=1) / argmax(abs2(x)) of complex<float> rows of 8192
    rng = np.random.default_rng(12)
    x = (rng.standard_normal((24, 8192)) + 1j * rng.standard_normal((24, 8192))).astype(np.complex64)  # zero mean, as config 3
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.mean(t, [1]), [x], A.C64)
    assert np.max(np.abs(got - want)) <= 1e-5 * np.mean(np.abs(x)), k   # norm-wise bound: the mean itself is ~0
    k = check(oracle, "var", lambda t: mx.var(t, [1], 1), [x], A.F32, tol=2e-5)
    assert k.startswith("var_tma"), k   # rows of 64 KB: TMA-staged ring in shared memory
    x2 = x[:, :1024].copy()             # rows of 8 KB: a warp per row, row in registers
    k = check(oracle, "var", lambda t: mx.var(t, [1], 1), [x2], A.F32, tol=2e-5)
    assert k.startswith("var_group"), k
    k = check(oracle, "var", lambda t: mx.var(t * 2.0, [1], 1), [x], A.F32, tol=2e-5)   # fused expression, 64 KB rows: CTA per row, registers
    assert k.startswith("var_reg"), k
    x3 = x[:, :48].copy()               # 48 complex per row: 12 vectors -> 16 lanes per row, 2 rows per warp
    k = check(oracle, "stdd", lambda t: mx.stdd(t, [1], 0), [x3], A.F32, tol=2e-5)
    assert k.startswith("var_group"), k
    got, gi, want, wi, k = G.run_reduce(oracle, lambda t: mx.argmax(mx.abs2(t), [1]), [x], A.F32)
    # abs2 contracts to an FMA on the device: values may differ in the last bit, the winner may not
    assert np.array_equal(gi, wi) and G.rel_err(got, want) < 1e-6, k


def test_var_two_launch_path(oracle, monkeypatch):
    rng = np.random.default_rng(13)
    x = (rng.random((3, 70000)) + 2).astype(np.float32)  # 280 KB rows: do not fit in shared memory
    k = check(oracle, "var", lambda t: mx.var(t, [1], 1), [x], A.F32, tol=5e-5)
    assert k.startswith("red_inner"), k
    y = data(rng, (50, 300), A.C64)
    monkeypatch.setenv("MXB_VAR_SMEM_ONLY", "1")
    k = check(oracle, "stdd", lambda t: mx.stdd(t, [1], 0), [y], A.F32, tol=2e-5)
    assert k.startswith("var_smem"), k
    monkeypatch.setenv("MXB_VAR_TWO_LAUNCH", "1")
    k = check(oracle, "stdd", lambda t: mx.stdd(t, [1], 0), [y], A.F32, tol=2e-5)
    assert k.startswith("red_inner"), k


def test_bf16_permuted_config5_shape(oracle):
    # config 5 at reduced size: sum(permute(t,{2,0,1}),{2}) on bf16 with fp32 accumulation, bf16 output
    rng = np.random.default_rng(14)
    f = (rng.random((16, 96, 256)) * 0.25).astype(np.float32)
    bits = f32_to_bf16_bits(f).reshape(f.shape)
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.sum(mx.permute(t, [2, 0, 1]), [2]), [bits], A.BF16, dtypes=[A.BF16])
    assert k.startswith("red_outer") and "|V8" in k, k
    g, w = bf16_bits_to_f32(got), bf16_bits_to_f32(want)
    assert g.shape == (256, 16)
    assert np.array_equal(got, want), float(np.max(np.abs(g - w) / w))   # fp32 accumulate + one rounding on both sides
    truth = bf16_bits_to_f32(bits).astype(np.float64).sum(axis=1).T
    assert np.max(np.abs(g - truth) / truth) <= 2 ** -8
    # secondary: {1,2}
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.sum(mx.permute(t, [2, 0, 1]), [1, 2]), [bits], A.BF16, dtypes=[A.BF16])
    assert G.rel_err(bf16_bits_to_f32(got), bf16_bits_to_f32(want)) <= 1e-2


def test_jit_arbitrary_expression_reduce(oracle):
    rng = np.random.default_rng(15)
    a = (rng.random((33, 500)) + 0.1).astype(np.float32)
    b = (rng.random((500,)) + 0.1).astype(np.float32)  # lower-rank operand broadcasts over rows
    k = check(oracle, "sum", lambda x, y: mx.sum(mx.sqrt(x) * y - mx.log(x + 1.0) / 3.0, [1]), [a, b], A.F32, tol=2e-5)
    assert k.endswith("jit"), k
    check(oracle, "argmax", lambda x, y: mx.argmax(mx.floor(x * 10.0) + mx.floor(y * 3.0), [1]), [a, b], A.F32)
    check(oracle, "any", lambda x, y: mx.any((x > 1.05) & (y > 0.5), [1]), [a, b], A.U8)


def test_error_convention():
    import torch
    ex = G.executor()
    x = mx.make_tensor(torch.zeros((4, 8), device="cuda"))
    bad = mx.make_tensor(torch.zeros((5,), device="cuda"))
    with pytest.raises(A.MatxB200Error) as ei:
        bad.set(mx.sum(x, [1]))          # matxInvalidSize at set construction (operators/set.h:196-198)
    assert ei.value.status == A.ERR_SIZE
    out = mx.make_tensor(torch.zeros((4,), device="cuda"))
    with pytest.raises(TypeError):
        out.set(mx.argmax(x, [1])).run(ex)
    c = mx.make_tensor(torch.zeros((4, 8), dtype=torch.complex64, device="cuda"))
    with pytest.raises(A.MatxB200Error) as ei:
        out.set(mx.max(c, [1])).run(ex)   # the reference rejects ordering of complex too
    assert ei.value.status == A.ERR_NOT_SUPPORTED


def test_broadcast_leaf_along_the_vector_dim(oracle):
    """A leaf whose stride along the vector dim is 0 (clone / per-row scalar) takes the splat path of the loads."""
    rng = np.random.default_rng(21)
    a = (rng.random((64, 1024)) + 0.5).astype(np.float32)
    rowv = (rng.random(64) + 0.5).astype(np.float32)       # one value per row, broadcast along the reduced dim
    colv = (rng.random(1024) + 0.5).astype(np.float32)     # one value per column, broadcast along the batch dim
    for op in ["sum", "max", "argmax", "var"]:
        f = getattr(mx, op)
        k = check(oracle, op, lambda x, r, c, f=f: f((x + c) * mx.clone(r, [mx.matxKeepDim, 1024]), [1]), [a, rowv, colv], A.F32, tol=2e-5)
        assert "|V4" in k, k                               # still the vector kernel; only the row scalar is splat
        k = check(oracle, op, lambda x, r, c, f=f: f((x + c) * mx.clone(r, [mx.matxKeepDim, 1024]), [0]), [a, rowv, colv], A.F32, tol=2e-5)
        assert k.startswith("red_outer") or op == "var", k
    got, want, k = G.run_elementwise(oracle, lambda x, r, c: x / mx.clone(r, [mx.matxKeepDim, 1024]) - c, [a, rowv, colv], a.shape, A.F32)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-6) and "|V4" in k, k


def test_handles_are_independent_and_streams_respected(oracle):
    import torch
    rng = np.random.default_rng(22)
    x = G.to_dev(rng.random((3, 1 << 20)).astype(np.float32))
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    e1, e2 = mx.CudaExecutor(s1), mx.CudaExecutor(s2)
    outs = []
    for ex_, st in ((e1, s1), (e2, s2)):
        with torch.cuda.stream(st):
            o = torch.zeros(3, device="cuda")
            for _ in range(20):                              # grid-combine scratch of the two handles must not interfere
                mx.make_tensor(o).set(mx.sum(mx.make_tensor(x), [1])).run(ex_)
            outs.append(o)
    e1.sync(); e2.sync()
    truth = x.double().sum(1)
    for o in outs:
        assert ((o.double() - truth).abs() / truth).max().item() <= 1e-5
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("cols", [1, 2, 3, 4, 7, 8, 20, 64, 100, 1000])
def test_short_rows_share_a_warp(oracle, cols):
    """Rows shorter than a warp's worth of vectors: G lanes per row, 32/G rows per warp (ragged last warp included)."""
    rng = np.random.default_rng(30 + cols)
    rows = 1237                                            # not a multiple of any rows-per-warp
    x = data(rng, (rows, cols), A.F32, ties=True)
    for op in ["sum", "max", "argmax", "argmin", "any", "all"]:
        check(oracle, op, lambda t, op=op: getattr(mx, op)(t, [1]), [x], A.F32)
    y = data(rng, (rows, cols), A.F32) + np.float32(0.5)
    check(oracle, "mean", lambda t: mx.mean(t, [1]), [y], A.F32)
    if cols >= 2:
        check(oracle, "var", lambda t: mx.var(t, [1], 1), [y], A.F32, tol=2e-5)


@pytest.mark.parametrize("shape", [(200_000, 8), (50_000, 33), (30_000, 128), (3_000, 1000)])
def test_column_reductions_of_tall_matrices(oracle, shape):
    """Few output tiles, long strided reduce dim: several CTAs per tile with the in-launch combine (reduce_outer splits)."""
    rng = np.random.default_rng(40 + shape[1])
    x = data(rng, shape, A.F32, ties=True)
    for op in ["max", "argmax", "argmin", "all"]:
        k = check(oracle, op, lambda t, op=op: getattr(mx, op)(t, [0]), [x], A.F32)
    assert k.startswith("red_outer"), k
    y = (data(rng, shape, A.F32) + np.float32(0.5))
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.sum(t, [0]), [y], A.F32)
    truth = y.astype(np.float64).sum(0)
    assert np.max(np.abs(got - truth) / truth) <= 1e-5, k   # the sequential fp32 oracle is the less accurate side here
    assert np.max(np.abs(want - truth) / truth) <= 5e-3


# ---- trace / diag / isclose / allclose (transforms/reduce.h:1321-1331,1505-1511; operators/diag.h, isclose.h) ----
def test_trace_and_diag(oracle):
    rng = np.random.default_rng(41)
    for n in (1, 7, 256, 1000):
        m = rng.standard_normal((n, n)).astype(np.float32)
        got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.trace(t), [m], A.F32)
        assert abs(float(got) - float(want)) <= 1e-5 * max(1.0, float(np.sum(np.abs(np.diag(m)))))
        assert abs(float(got) - float(np.trace(m.astype(np.float64)))) <= 1e-5 * max(1.0, float(np.sum(np.abs(np.diag(m)))))
    # batched: trace of every matrix of a stack, and of a fused expression
    s = rng.standard_normal((9, 33, 33)).astype(np.float32)
    got, _, want, _, _ = G.run_reduce(oracle, lambda t: mx.sum(mx.diag(t * t), [1]), [s], A.F32)
    ref = np.einsum("bii->b", s.astype(np.float64) ** 2)
    assert np.allclose(got, want, rtol=1e-5) and np.allclose(got, ref, rtol=1e-5)
    # off-diagonals of a square matrix as views
    m = rng.standard_normal((64, 64)).astype(np.float32)
    for kk in (1, -3):
        got, want, _ = G.run_elementwise(oracle, lambda t: mx.diag(t, kk) * 1.0, [m], (64 - abs(kk),), A.F32)
        assert np.array_equal(got, np.diag(m, kk)) and np.array_equal(got, want)
    with pytest.raises(ValueError):
        mx.diag(mx.make_tensor(__import__("torch").zeros(7, 5, device="cuda")), 1)


def test_isclose_allclose(oracle):
    import torch
    rng = np.random.default_rng(42)
    a = rng.standard_normal((300, 70)).astype(np.float32)
    b = a.copy()
    b[17, 3] += 1e-3
    got, want, _ = G.run_elementwise(oracle, lambda x, y: mx.isclose(x, y, 1e-5, 1e-8), [a, b], a.shape, A.I32)
    assert np.array_equal(got, want)
    assert np.array_equal(got.astype(bool), np.abs(a - b) <= np.float32(1e-8) + np.float32(1e-5) * np.abs(b))
    ex = G.executor()
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    flag = torch.full((), -1, dtype=torch.int32, device="cuda")
    mx.allclose(mx.make_tensor(flag), mx.make_tensor(da), mx.make_tensor(db), 1e-5, 1e-8, ex)
    ex.sync()
    assert int(flag) == 0
    mx.allclose(mx.make_tensor(flag), mx.make_tensor(da), mx.make_tensor(db), 1e-2, 1e-8, ex)
    ex.sync()
    assert int(flag) == 1
    mx.allclose(mx.make_tensor(flag), mx.make_tensor(da), mx.make_tensor(da), 0.0, 0.0, ex)
    ex.sync()
    assert int(flag) == 1
    with pytest.raises(TypeError):
        mx.allclose(mx.make_tensor(torch.zeros(3, dtype=torch.int32, device="cuda")), mx.make_tensor(da), mx.make_tensor(db), 1e-5, 1e-8, ex)
