"""CPU-side tests of the boundary: the C-ABI library loads without a GPU, exports every symbol include/matx_b200.h
declares, refuses to compute without a device (no CPU fallback), and the host-side lowering / code generator behave.
NVRTC is used in compile-only mode to prove that generated kernels build for sm_100a on a box with no GPU."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests.oracle_harness import np_tensor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(built_lib):
    hdr = open(os.path.join(ROOT, "include", "matx_b200.h")).read()
    declared = set(re.findall(r"\b(mxb_[a-z_]+)\s*\(", hdr))
    declared -= {"mxb_context"}
    assert declared == set(A.EXPORTED), declared ^ set(A.EXPORTED)
    lib = C.CDLL(built_lib)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.mxb_version() == 1


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    assert A.lib.mxb_device_count() == 0
    h = C.c_void_p()
    st = A.lib.mxb_create(C.byref(h), None)
    assert st == A.ERR_NO_DEVICE and not h.value
    assert b"no CPU fallback" in A.lib.mxb_last_error()
    with pytest.raises(A.MatxB200Error) as ei:
        mx.CudaExecutor()
    assert ei.value.status == A.ERR_NO_DEVICE
    # compute entry points reject a null handle instead of computing anything
    x = np_tensor(np.ones(8, np.float32))
    e = mx.lower_reduce(mx.sum(x))
    o = mx._out_desc(np_tensor(np.zeros((), np.float32)))
    assert A.lib.mxb_reduce(None, A.RED_SUM, C.byref(e), 1, C.byref(o), None, 1) == A.ERR_INVALID
    assert A.lib.mxb_elementwise(None, C.byref(mx.lower_elementwise(x + 1.0)), C.byref(mx._out_desc(x))) == A.ERR_INVALID


def codegen(expr) -> str:
    buf = C.create_string_buffer(1 << 16)
    A.check(A.lib.mxb_debug_codegen(C.byref(expr), buf, len(buf)))
    return buf.value.decode()


def test_lowering_of_sum_dims_moves_reduced_dims_last():
    # sum(x, {0}) on a contiguous 4x8x16 tensor == trailing-dim reduce of the view with strides (16, 1, 128)
    # (getPermuteDims, core/utils.h:96-127; SURVEY.md section 8a row a6 states the reference's result on this shape)
    t = np_tensor(np.zeros((4, 8, 16), np.float32))
    r = mx.sum(mx.permute(t, [2, 0, 1]), [2])      # config 5 statement
    e = mx.lower_reduce(r)
    assert r.out_shape == (16, 4) and e.rank == 3
    assert list(e.size[:3]) == [16, 4, 8] and list(e.leaves[0].stride[:3]) == [1, 128, 16]
    r = mx.sum(t, [0])
    e = mx.lower_reduce(r)
    assert list(e.size[:3]) == [8, 16, 4] and list(e.leaves[0].stride[:3]) == [16, 1, 128]


def test_broadcast_clone_and_cse():
    a = np_tensor(np.zeros((5, 7), np.float32))
    v = np_tensor(np.zeros(7, np.float32))
    e = mx.lower_elementwise(a * v + mx.clone(v, [5, mx.matxKeepDim]) + a)
    src = codegen(e)
    assert src.count("mxb::ldleaf<") == 2            # a and v are each loaded once although they appear twice
    assert list(e.leaves[1].stride[:2]) == [0, 1]    # lower-rank operand lines up with the trailing dim


def test_types_follow_cpp_promotion():
    f = np_tensor(np.zeros(4, np.float32))
    d = np_tensor(np.zeros(4, np.float64))
    i = np_tensor(np.zeros(4, np.int32))
    c = np_tensor(np.zeros(4, np.complex64))
    assert "value_dtype=f64" in codegen(mx.lower_elementwise(f + d))
    assert "value_dtype=f32" in codegen(mx.lower_elementwise(f * i))
    assert "value_dtype=i32" in codegen(mx.lower_elementwise(i / i))
    assert "value_dtype=f64" in codegen(mx.lower_elementwise(mx.sqrt(i)))
    assert "value_dtype=c64" in codegen(mx.lower_elementwise(c * f))
    assert "value_dtype=f32" in codegen(mx.lower_elementwise(mx.abs2(c)))
    assert "value_dtype=u8" in codegen(mx.lower_elementwise(f > 0.5))
    with pytest.raises(A.MatxB200Error) as ei:
        codegen(mx.lower_elementwise(c * d))           # complex<double> is not lowered
    assert ei.value.status == A.ERR_NOT_SUPPORTED


def test_named_programs_are_ahead_of_time():
    f = lambda *s: np_tensor(np.zeros(s, np.float32))  # noqa: E731
    a, b, c = f(4, 8), f(4, 8), f(4, 8)
    assert A.lib.mxb_is_aot(C.byref(mx.lower_reduce(mx.sum(a * b + c, [1]))), A.RED_SUM) == 1
    assert A.lib.mxb_is_aot(C.byref(mx.lower_reduce(mx.argmax(a))), A.RED_ARGMAX) == 1
    x = np_tensor(np.zeros((4, 8), np.complex64))
    assert A.lib.mxb_is_aot(C.byref(mx.lower_reduce(mx.argmax(mx.abs2(x), [1]))), A.RED_ARGMAX) == 1
    K, S, V, r, T = (f(16) for _ in range(5))
    VsqrtT = V * mx.sqrt(T)
    d1 = (mx.log(S / K) + (r + 0.5 * V * V) * T) / VsqrtT
    d2 = d1 - VsqrtT
    bs = S * mx.normcdf(d1) - K * mx.exp(-1.0 * r * T) * mx.normcdf(d2)
    assert A.lib.mxb_is_aot(C.byref(mx.lower_elementwise(bs)), -1) == 1
    assert A.lib.mxb_is_aot(C.byref(mx.lower_elementwise(mx.tanh(a) * 3.0)), -1) == 0


@pytest.mark.parametrize("family,op,team", [(0, A.RED_SUM, 0), (0, A.RED_ARGMIN, 1), (1, A.RED_MAX, 0), (2, A.RED_VAR, 0),
                                            (4, A.RED_VAR, 4), (3, -1, 0), (6, A.RED_VAR, 2), (7, -1, 2), (8, -1, 4),
                                            (0, 11, 0), (1, 11, 0), (9, -1, 0), (10, -1, 0)])   # 7 / 8: softmax in registers; op 11: its statistics pass; 9: transposing elementwise; 10: scan
def test_generated_kernels_compile_for_sm100a(family, op, team):
    a = np_tensor(np.zeros((4, 8), np.float32))
    h = np_tensor(np.zeros((4, 8), np.uint16), A.BF16)
    expr = mx.sqrt(mx.abs(a)) * h - mx.tanh(a) / 3.0 + mx.as_type(mx.floor(a), A.I32)
    e = mx.lower_reduce(mx.ReduceExpr(max(op, 0), expr, [1])) if family not in (3, 9, 10) else mx.lower_elementwise(expr)
    log = C.create_string_buffer(1 << 16)
    st = A.lib.mxb_debug_compile(C.byref(e), family, op, A.F32, 0, team, log, len(log))
    assert st == A.OK, (A.lib.mxb_last_error(), log.value.decode()[:2000])


@pytest.mark.parametrize("npdt,dt,op", [(np.int32, A.I32, A.RED_SUM), (np.int32, A.I32, A.RED_ARGMIN), (np.float64, A.F64, A.RED_MAX),
                                        (np.complex64, A.C64, A.RED_SUM), (np.float32, A.F32, A.RED_ALL)])
def test_outer_tma_kernel_compiles_for_sm100a(npdt, dt, op):
    """Family 11 (red_outer_tma: TMA-staged tiles for a strided reduce dim) serves plain tensors; the dtypes without an
    ahead-of-time instance are built by NVRTC on first use — prove here that they compile (V = one 16-byte chunk)."""
    x = np_tensor(np.zeros((8, 16), npdt))
    e = mx.lower_reduce(mx.ReduceExpr(op, x, [0]))
    log = C.create_string_buffer(1 << 16)
    st = A.lib.mxb_debug_compile(C.byref(e), 11, op, dt, 16 // np.dtype(npdt).itemsize, 0, log, len(log))
    assert st == A.OK, (A.lib.mxb_last_error(), log.value.decode()[:2000])


@pytest.mark.parametrize("team,out_dt", [(0, A.F32), (1, A.F32), (2, A.I32), (2, A.I64), (3, A.F32), (3, A.F64), (4, A.I32), (4, A.I64)])
def test_select_kernels_compile_for_sm100a(team, out_dt):
    """Family 12 (find / find_idx): count pass, value scatter, index scatter — of an expression operand, vector and scalar;
    teams 3 / 4 are the single-pass look-back kernel (values / flat indices)."""
    a = np_tensor(np.zeros(64, np.float32))
    e = mx.lower_elementwise(mx.sqrt(mx.abs(a)) * 2.0 - 1.0)
    for V in (4, 1):
        log = C.create_string_buffer(1 << 16)
        st = A.lib.mxb_debug_compile(C.byref(e), 12, -1, out_dt, V, team, log, len(log))
        assert st == A.OK, (A.lib.mxb_last_error(), log.value.decode()[:2000])
    assert A.lib.mxb_find(None, C.byref(e), 0, 0.0, None, None, 0) == A.ERR_INVALID   # null handle: an error, not a crash


def test_paired_fp32_body_compiles_for_sm100a():
    """Pure-fp32 programs get a second body on the packed fp32 instructions (FFMA2 / FMUL2 / FADD2): NVRTC must know
    the sm_100 intrinsics and the packed log / normcdf."""
    a, b = (np_tensor(np.zeros((4, 8), np.float32)) for _ in range(2))
    expr = mx.normcdf(mx.log(a / b) * 2.0 - mx.sqrt(b)) * a - mx.exp(-a) + mx.abs(b)
    e = mx.lower_elementwise(expr)
    buf = C.create_string_buffer(1 << 16)
    assert A.lib.mxb_debug_codegen(C.byref(e), buf, len(buf)) == A.OK
    src = buf.value.decode()
    assert "PAIR = 1" in src and "eval2" in src and "mxb::fma2(" in src, src
    log = C.create_string_buffer(1 << 16)
    st = A.lib.mxb_debug_compile(C.byref(e), 3, -1, A.F32, 0, 0, log, len(log))
    assert st == A.OK, (A.lib.mxb_last_error(), log.value.decode()[:2000])
    # anything outside the packed subset keeps the scalar body only
    e2 = mx.lower_elementwise(mx.tanh(a) + b)
    assert A.lib.mxb_debug_codegen(C.byref(e2), buf, len(buf)) == A.OK
    assert "PAIR = 0" in buf.value.decode()


def test_set_checks_shapes_like_the_reference():
    a = np_tensor(np.zeros((4, 8), np.float32))
    with pytest.raises(A.MatxB200Error) as ei:
        np_tensor(np.zeros(5, np.float32)).set(mx.sum(a, [1]))
    assert ei.value.status == A.ERR_SIZE
    with pytest.raises(ValueError):
        a + np_tensor(np.zeros(7, np.float32))
    with pytest.raises(TypeError):
        mx.mtie(a, a).set(mx.sum(a))


def test_malformed_programs_are_rejected_not_crashed():
    """A cyclic program (a -> b -> a) or a rank outside [0, MXB_MAX_RANK] must come back as MXB_ERR_INVALID from every
    entry point that canonicalises (the cycle used to exhaust memory, the rank to read out of bounds)."""
    x = np_tensor(np.zeros(8, np.float32))
    e = mx.lower_elementwise(-x)
    buf = C.create_string_buffer(4096)
    cyc = A.Expr.from_buffer_copy(e)
    cyc.n_nodes = 2
    cyc.nodes[0].opcode, cyc.nodes[0].src[0] = A.OP_NEG, 1
    cyc.nodes[1].opcode, cyc.nodes[1].src[0] = A.OP_NEG, 0
    cyc.root = 0
    assert A.lib.mxb_debug_codegen(C.byref(cyc), buf, len(buf)) == A.ERR_INVALID
    assert b"cyclic" in A.lib.mxb_last_error()
    assert A.lib.mxb_is_aot(C.byref(cyc), -1) == 0
    for bad_rank in (-1, A.MXB_MAX_RANK + 1, 1 << 20):
        r = A.Expr.from_buffer_copy(e)
        r.rank = bad_rank
        assert A.lib.mxb_debug_codegen(C.byref(r), buf, len(buf)) == A.ERR_INVALID, bad_rank
        assert A.lib.mxb_is_aot(C.byref(r), -1) == 0
        log = C.create_string_buffer(256)
        assert A.lib.mxb_debug_compile(C.byref(r), 3, -1, A.F32, 1, 0, log, len(log)) == A.ERR_INVALID
