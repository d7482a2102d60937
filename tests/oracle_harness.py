"""Test-side access to the checkers: the C restatement (oracle/matx_oracle.c) and, when present, the compiled
reference (oracle/_ref/).  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline import this."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from matx_b200 import _abi as A  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

_NP2MXB = {np.dtype(np.float32): A.F32, np.dtype(np.float64): A.F64, np.dtype(np.complex64): A.C64,
           np.dtype(np.int32): A.I32, np.dtype(np.int64): A.I64, np.dtype(np.uint8): A.U8, np.dtype(np.bool_): A.U8}


def np_tensor(arr: np.ndarray, dtype: int | None = None) -> mx.Tensor:
    """View a numpy array (any strides) as an expression leaf.  16-bit floats travel as uint16 + dtype."""
    if dtype is None:
        dtype = _NP2MXB[arr.dtype]
    item = arr.dtype.itemsize
    strides = [s // item for s in arr.strides]
    return mx.Tensor(arr.ctypes.data, dtype, arr.shape, strides, keepalive=arr)


def f32_to_bf16_bits(x: np.ndarray) -> np.ndarray:
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    r = (u + 0x7FFF + ((u >> 16) & 1)) >> 16
    return r.astype(np.uint16)


def bf16_bits_to_f32(b: np.ndarray) -> np.ndarray:
    return (b.astype(np.uint32) << 16).view(np.float32)


class Oracle:
    def __init__(self, lib: C.CDLL):
        self.lib = lib
        lib.orc_elementwise.argtypes = [C.POINTER(A.Expr), C.POINTER(A.Out)]
        lib.orc_reduce.argtypes = [C.c_int, C.POINTER(A.Expr), C.c_int, C.POINTER(A.Out), C.POINTER(A.Out), C.c_int, C.c_int]
        lib.orc_softmax.argtypes = [C.POINTER(A.Expr), C.c_int, C.POINTER(A.Out)]
        lib.orc_cumsum.argtypes = [C.POINTER(A.Expr), C.POINTER(A.Out)]
        lib.orc_find.argtypes = [C.POINTER(A.Expr), C.c_int, C.c_double, C.POINTER(A.Out), C.POINTER(C.c_int32), C.c_int]
        lib.orc_sort.argtypes = [C.POINTER(A.Expr), C.POINTER(A.Out), C.c_int]
        lib.orc_unique.argtypes = [C.POINTER(A.Expr), C.POINTER(A.Out), C.POINTER(C.c_int32)]
        lib.orc_hist.argtypes = [C.POINTER(A.Expr), C.c_double, C.c_double, C.POINTER(A.Out)]

    def elementwise(self, rhs, out: np.ndarray, out_dtype: int | None = None) -> np.ndarray:
        rhs = mx._wrap(rhs, None)
        lhs = np_tensor(out, out_dtype)
        if len(rhs.shape) < len(lhs.shape):
            rhs = mx.CloneOp(rhs, list(lhs.shape[:len(lhs.shape) - len(rhs.shape)]) + [mx.matxKeepDim] * len(rhs.shape))
        e = mx.lower_elementwise(rhs)
        o = mx._out_desc(lhs)
        assert self.lib.orc_elementwise(C.byref(e), C.byref(o)) == 0
        return out

    def find(self, r: "mx.FindExpr", out: np.ndarray) -> int:
        """Fills `out` (rank 1) and returns num_found."""
        e = mx.lower_elementwise(r.a)
        o = mx._out_desc(np_tensor(out))
        n = C.c_int32(0)
        assert self.lib.orc_find(C.byref(e), r.sel.op, float(r.sel.c), C.byref(o), C.byref(n), 1 if r.want_indices else 0) == 0
        return int(n.value)

    def sort(self, r: "mx.SortExpr", out: np.ndarray) -> np.ndarray:
        e = mx.lower_elementwise(r.a)
        o = mx._out_desc(np_tensor(out))
        assert self.lib.orc_sort(C.byref(e), C.byref(o), 1 if r.direction == mx.SORT_DIR_DESC else 0) == 0
        return out

    def unique(self, r: "mx.UniqueExpr", out: np.ndarray) -> int:
        e = mx.lower_elementwise(r.a)
        o = mx._out_desc(np_tensor(out))
        n = C.c_int32(0)
        assert self.lib.orc_unique(C.byref(e), C.byref(o), C.byref(n)) == 0
        return int(n.value)

    def hist(self, r: "mx.HistExpr", out: np.ndarray) -> np.ndarray:
        e = mx.lower_elementwise(r.a)
        o = mx._out_desc(np_tensor(out))
        assert self.lib.orc_hist(C.byref(e), float(r.lower), float(r.upper), C.byref(o)) == 0
        return out

    def cumsum(self, r: "mx.CumsumExpr", out: np.ndarray, out_dtype: int | None = None) -> np.ndarray:
        e = mx.lower_elementwise(r.a)
        o = mx._out_desc(np_tensor(out, out_dtype))
        assert self.lib.orc_cumsum(C.byref(e), C.byref(o)) == 0
        return out

    def softmax(self, r: "mx.SoftmaxExpr", out: np.ndarray, out_dtype: int | None = None) -> np.ndarray:
        e = mx.lower_reduce(r)
        o = mx._out_desc(np_tensor(out, out_dtype))
        po = A.Out()
        po.data, po.dtype, po.rank = o.data, o.dtype, o.rank
        for i, d in enumerate(r.perm):
            po.size[i], po.stride[i] = o.size[d], o.stride[d]
        assert self.lib.orc_softmax(C.byref(e), len(r.dims), C.byref(po)) == 0
        return out

    def reduce(self, r: mx.ReduceExpr, out: np.ndarray, idx: np.ndarray | None = None, out_dtype: int | None = None,
               half_acc: int = -1) -> None:
        e = mx.lower_reduce(r)
        o = mx._out_desc(np_tensor(out, out_dtype))
        io = mx._out_desc(np_tensor(idx)) if idx is not None else None
        rc = self.lib.orc_reduce(r.op, C.byref(e), len(r.dims), C.byref(o), C.byref(io) if io is not None else None, r.ddof,
                                 half_acc)
        assert rc == 0


def load_oracle() -> Oracle:
    path = build_ref.build_oracle()
    return Oracle(C.CDLL(path))


# ---- the compiled reference (oracle/_ref), optional ------------------------------------------------------
_REF_DT = {A.F32: 0, A.F64: 1, A.BF16: 2, A.C64: 4, A.I32: 5}


class RefHost:
    """matx::HostExecutor statements compiled from /root/reference (oracle/ref_wrap.cu)."""

    def __init__(self, lib: C.CDLL, suffix: str = "_host"):
        self.lib = lib
        self.suffix = suffix

    def fn(self, name):
        return getattr(self.lib, name + self.suffix)

    def reduce(self, op: int, arr_ptr: int, dtype: int, shape, strides, dims, out_ptr: int, idx_ptr: int | None, ddof: int = 1,
               mode: int = 0) -> int:
        f = self.fn("mref_reduce_dt%d_" % _REF_DT[dtype])
        rank = len(shape)
        sh = (C.c_int64 * rank)(*shape)
        st = (C.c_int64 * rank)(*strides)
        dm = (C.c_int * len(dims))(*dims)
        f.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.c_void_p, C.c_int,
                      C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int]
        return f(mode, op, rank, sh, st, C.c_void_p(arr_ptr), len(dims), dm, C.c_void_p(out_ptr),
                 C.c_void_p(idx_ptr) if idx_ptr else None, ddof)

    def reduce_np(self, op: int, arr: np.ndarray, dims, ddof: int = 1, mode: int = 0, dtype: int | None = None):
        """Reduce `dims` of a (possibly strided) numpy view; returns (values, indices-or-None)."""
        dtype = _NP2MXB[arr.dtype] if dtype is None else dtype
        item = arr.dtype.itemsize
        out_shape = tuple(s for d, s in enumerate(arr.shape) if d not in dims)
        if op in (A.RED_VAR, A.RED_STDD) and dtype == A.C64:
            out = np.zeros(out_shape, np.float32)
        else:
            out = np.zeros(out_shape, arr.dtype)
        idx = np.zeros(out_shape, np.int64) if op in (A.RED_ARGMAX, A.RED_ARGMIN) else None
        rc = self.reduce(op, arr.ctypes.data, dtype, arr.shape, [s // item for s in arr.strides], list(dims), out.ctypes.data,
                         idx.ctypes.data if idx is not None else None, ddof, mode)
        if rc != 0:
            raise RuntimeError("reference wrapper returned %d" % rc)
        return out, idx


def load_ref_host() -> RefHost | None:
    path = build_ref.build_ref(cuda=False)
    if not path or not os.path.exists(path):
        return None
    return RefHost(C.CDLL(path))


def load_ref_cuda() -> RefHost | None:
    path = build_ref.REF_CUDA_LIB
    if not os.path.exists(path):
        return None
    return RefHost(C.CDLL(path), "_cuda")
