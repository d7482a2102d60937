"""Host-side mirror of the MatX operator surface for the reduce / fused-elementwise path.

The real drop-in is the C++ header shim (include/matx_b200/executor.h) that sits under the unmodified MatX
headers.  This module is the same lowering written in Python so that the parity tests and bench.py can state
MatX statements and send them through the same C ABI:

    MatX (C++)                                        here
    (out = sum(a*b+c, {1})).run(exec);                out.set(sum(a*b+c, [1])).run(exec)
    (mtie(v, i) = argmax(x, {1})).run(exec);          mtie(v, i).set(argmax(x, [1])).run(exec)
    (o = S*normcdf(d1) - K*exp(-1.f*r*T)*...).run()   o.set(S*normcdf(d1) - K*exp(-1.0*r*T)*...).run(exec)
    permute(t, {2,0,1})  /  t.Permute({2,0,1})        permute(t, [2,0,1])  /  t.Permute([2,0,1])
    clone<2>(v, {matxKeepDim, N})                     clone(v, [matxKeepDim, N])

Reference: operators/base_operator.h:181-285 (run dispatch), operators/set.h:121-553, core/tie.h:44-117,
operators/binary_operators.h:91-381, operators/unary_operators.h, operators/permute.h:389-397,
operators/clone.h, operators/sum.h:373-405 (dims -> permute + trailing-dim reduce), executors/cuda.h:60-82.
torch is used for device memory and streams only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _abi as A

import builtins

matxKeepDim = -1  # clone(): keep this input dim
builtins_min = builtins.min  # this module defines its own min / max / abs / sum / any / all (the MatX names)

_TORCH_DTYPES = None


def _torch_dtype_map():
    global _TORCH_DTYPES
    if _TORCH_DTYPES is None:
        import torch
        _TORCH_DTYPES = {
            torch.float32: A.F32, torch.float64: A.F64, torch.bfloat16: A.BF16, torch.float16: A.F16,
            torch.complex64: A.C64, torch.int32: A.I32, torch.int64: A.I64, torch.uint8: A.U8, torch.bool: A.U8,
        }
    return _TORCH_DTYPES


def _is_float_dtype(d: int) -> bool:
    return d in (A.F32, A.F64, A.BF16, A.F16)


# --------------------------------------------------------------------------------------------------------
# expression nodes
# --------------------------------------------------------------------------------------------------------
class Op:
    """Lazy operator node (reference: BaseOp<T>, operators/base_operator.h)."""
    shape: tuple
    dtype_hint: int  # mxb dtype used to type Python scalars that meet this operand

    def Rank(self) -> int:
        return len(self.shape)

    def Size(self, d: int) -> int:
        return self.shape[d]

    # arithmetic
    def __add__(self, o): return BinOp(A.OP_ADD, self, o)
    def __radd__(self, o): return BinOp(A.OP_ADD, o, self)
    def __sub__(self, o): return BinOp(A.OP_SUB, self, o)
    def __rsub__(self, o): return BinOp(A.OP_SUB, o, self)
    def __mul__(self, o): return BinOp(A.OP_MUL, self, o)
    def __rmul__(self, o): return BinOp(A.OP_MUL, o, self)
    def __truediv__(self, o): return BinOp(A.OP_DIV, self, o)
    def __rtruediv__(self, o): return BinOp(A.OP_DIV, o, self)
    def __mod__(self, o): return BinOp(A.OP_MOD, self, o)
    def __pow__(self, o): return BinOp(A.OP_POW, self, o)
    def __neg__(self): return UnOp(A.OP_NEG, self)
    # comparisons / logic (elementwise, like MatX)
    def __lt__(self, o): return BinOp(A.OP_LT, self, o)
    def __gt__(self, o): return BinOp(A.OP_GT, self, o)
    def __le__(self, o): return BinOp(A.OP_LE, self, o)
    def __ge__(self, o): return BinOp(A.OP_GE, self, o)
    def eq(self, o): return BinOp(A.OP_EQ, self, o)
    def ne(self, o): return BinOp(A.OP_NE, self, o)
    def __and__(self, o): return BinOp(A.OP_AND, self, o)
    def __or__(self, o): return BinOp(A.OP_OR, self, o)
    def __invert__(self): return UnOp(A.OP_NOT, self)
    __hash__ = object.__hash__


class Tensor(Op):
    """Non-owning strided view of device memory (reference: tensor_impl_t, core/tensor_impl.h)."""

    def __init__(self, data_ptr: int, dtype: int, shape: Sequence[int], strides: Sequence[int], keepalive=None):
        self.data_ptr = int(data_ptr)
        self.dtype = int(dtype)
        self.dtype_hint = self.dtype
        self.shape = tuple(int(s) for s in shape)
        self.strides = tuple(int(s) for s in strides)
        self._keep = keepalive
        if len(self.shape) != len(self.strides):
            raise ValueError("shape / strides rank mismatch")
        if len(self.shape) > A.MXB_MAX_RANK:
            raise ValueError("rank > %d" % A.MXB_MAX_RANK)

    # MatX spelling
    def Stride(self, d: int) -> int:
        return self.strides[d]

    def Data(self) -> int:
        return self.data_ptr

    def IsContiguous(self) -> bool:
        w = 1
        for s, st in zip(reversed(self.shape), reversed(self.strides)):
            if s != 1 and st != w:
                return False
            w *= s
        return True

    def Permute(self, dims: Sequence[int]) -> "Tensor":
        """tensor_t::Permute (core/tensor.h:894-900): a strided view, no data movement."""
        dims = list(dims)
        if sorted(dims) != list(range(len(self.shape))):
            raise ValueError("Permute: dims must be a permutation of 0..rank-1")
        return Tensor(self.data_ptr, self.dtype, [self.shape[d] for d in dims], [self.strides[d] for d in dims], self._keep)

    def Slice(self, starts: Sequence[int], ends: Sequence[int]) -> "Tensor":
        """Half-open slice per dim (matxEnd == None keeps the dim's end); rank is preserved."""
        off, shape = 0, []
        for d, (s, e) in enumerate(zip(starts, ends)):
            e = self.shape[d] if e is None else e
            if not (0 <= s <= e <= self.shape[d]):
                raise ValueError("Slice: range outside the tensor in dim %d" % d)
            off += s * self.strides[d]
            shape.append(e - s)
        return Tensor(self.data_ptr + off * A.DTYPE_BYTES[self.dtype], self.dtype, shape, self.strides, self._keep)

    def set(self, rhs) -> "Set":
        """`(self = rhs)` — the assignment node (operators/set.h)."""
        return Set(self, rhs)


def make_tensor(t) -> Tensor:
    """Wrap a torch tensor (any strides) without copying — make_tensor(T* data, shape, strides),
    core/make_tensor.h:710-."""
    dt = _torch_dtype_map().get(t.dtype)
    if dt is None:
        raise TypeError("unsupported dtype %s" % t.dtype)
    return Tensor(t.data_ptr(), dt, tuple(t.shape), tuple(t.stride()), keepalive=t)


class Const(Op):
    def __init__(self, value, dtype: int):
        self.value = value
        self.dtype = dtype
        self.dtype_hint = dtype
        self.shape = ()


def _wrap(x, other: Optional[Op]) -> Op:
    if isinstance(x, Op):
        return x
    hint = other.dtype_hint if other is not None else A.F32
    if isinstance(x, bool):
        return Const(int(x), A.U8)
    if isinstance(x, complex):
        return Const(x, A.C64)
    if isinstance(x, int):
        # `static_cast<inner_type>(2)` idiom: an integer literal takes the floating type it meets
        if _is_float_dtype(hint) or hint == A.C64:
            return Const(float(x), A.F32 if hint in (A.BF16, A.F16, A.C64) else hint)
        return Const(x, hint if hint in (A.I32, A.I64) else A.I32)
    if isinstance(x, float):
        if hint == A.F64:
            return Const(x, A.F64)
        return Const(x, A.F32)
    raise TypeError("cannot use %r in a matx_b200 expression" % (x,))


def _broadcast_shape(a: tuple, b: tuple) -> tuple:
    """MatX rule (MATX_ASSERT_COMPATIBLE_OP_SIZES): a lower-rank operand lines up with the trailing dims;
    sizes must match (rank-0 operands broadcast everywhere)."""
    if len(a) < len(b):
        a, b = b, a
    off = len(a) - len(b)
    for i, s in enumerate(b):
        if s != a[off + i]:
            raise ValueError("incompatible operator sizes %s vs %s" % (a, b))
    return a


class BinOp(Op):
    def __init__(self, opcode: int, a, b):
        a = _wrap(a, b if isinstance(b, Op) else None)
        b = _wrap(b, a)
        self.opcode, self.a, self.b = opcode, a, b
        self.shape = _broadcast_shape(a.shape, b.shape)
        ha, hb = a.dtype_hint, b.dtype_hint
        rank = {A.U8: 0, A.I32: 1, A.I64: 2, A.BF16: 3, A.F16: 3, A.F32: 4, A.F64: 5, A.C64: 6}
        self.dtype_hint = ha if rank[ha] >= rank[hb] else hb


class UnOp(Op):
    def __init__(self, opcode: int, a, aux: int = 0):
        a = _wrap(a, None)
        self.opcode, self.a, self.aux = opcode, a, aux
        self.shape = a.shape
        h = a.dtype_hint
        if opcode in (A.OP_ABS, A.OP_ABS2, A.OP_REAL, A.OP_IMAG) and h == A.C64:
            h = A.F32
        if opcode == A.OP_CAST:
            h = aux
        if opcode == A.OP_EXPJ:
            h = A.C64
        self.dtype_hint = h


class PermuteOp(Op):
    """permute(op, dims) on a non-tensor operand (operators/permute.h:389-397)."""

    def __init__(self, a: Op, dims: Sequence[int]):
        dims = list(dims)
        if sorted(dims) != list(range(len(a.shape))):
            raise ValueError("permute: dims must be a permutation of 0..rank-1")
        self.a, self.dims = a, dims
        self.shape = tuple(a.shape[d] for d in dims)
        self.dtype_hint = a.dtype_hint


class CloneOp(Op):
    """clone<N>(op, {…}) — broadcast along new dims (operators/clone.h)."""

    def __init__(self, a: Op, cdims: Sequence[int]):
        keep = [i for i, c in enumerate(cdims) if c == matxKeepDim]
        if len(keep) != len(a.shape):
            raise ValueError("clone: number of matxKeepDim entries must equal the operand rank")
        self.a, self.cdims = a, list(cdims)
        it = iter(a.shape)
        self.shape = tuple(next(it) if c == matxKeepDim else int(c) for c in cdims)
        self.dtype_hint = a.dtype_hint


def permute(a, dims):
    return a.Permute(dims) if isinstance(a, Tensor) else PermuteOp(a, dims)


def clone(a, cdims):
    return CloneOp(a, cdims)


def _un(opcode):
    def f(a):
        return UnOp(opcode, a)
    return f


sqrt, rsqrt, exp, log, log2, log10 = (_un(o) for o in (A.OP_SQRT, A.OP_RSQRT, A.OP_EXP, A.OP_LOG, A.OP_LOG2, A.OP_LOG10))
abs, abs2, conj, real, imag = (_un(o) for o in (A.OP_ABS, A.OP_ABS2, A.OP_CONJ, A.OP_REAL, A.OP_IMAG))  # noqa: A001
sin, cos, tan, tanh, sinh, cosh = (_un(o) for o in (A.OP_SIN, A.OP_COS, A.OP_TAN, A.OP_TANH, A.OP_SINH, A.OP_COSH))
asin, acos, atan = (_un(o) for o in (A.OP_ASIN, A.OP_ACOS, A.OP_ATAN))
normcdf, isnan, isinf, floor, ceil, expj = (_un(o) for o in (A.OP_NORMCDF, A.OP_ISNAN, A.OP_ISINF, A.OP_FLOOR, A.OP_CEIL, A.OP_EXPJ))
round_ = _un(A.OP_ROUND)


class DiagOp(Op):
    """diag(op, k) of a rank >= 2 operand (operators/diag.h:145-190): the k-th diagonal of the last two dims; the result
    has one dim fewer, its last dim walks rows and columns together (one stride = s_row + s_col)."""

    def __init__(self, a: Op, k: int = 0):
        if len(a.shape) < 2:
            raise ValueError("diag: operand rank must be >= 2 (building a matrix from a vector is not on this path)")
        rows, cols = a.shape[-2], a.shape[-1]
        # the reference sizes the result as min(cols, rows - k) / min(cols + k, rows) (operators/diag.h:253-268) while it
        # INDEXES (i, i + k) / (i - k, i); the two agree for square matrices.  Where its size would run off the matrix
        # (non-square, k != 0) this mirror refuses instead of reading out of bounds.
        valid = builtins_min(rows, cols - k) if k >= 0 else builtins_min(rows + k, cols)
        n = builtins_min(cols, rows - k) if k > 0 else (builtins_min(cols + k, rows) if k < 0 else valid)
        if n <= 0 or n > valid:
            raise ValueError("diag: diagonal %d of a %dx%d matrix is outside it (or the reference's size rule is)" % (k, rows, cols))
        if k != 0 and not isinstance(a, Tensor):
            raise A.MatxB200Error(A.ERR_NOT_SUPPORTED, "diag(expression, k != 0): take the diagonal of the tensors instead")
        self.a, self.k = a, int(k)
        self.shape = tuple(a.shape[:-2]) + (n,)
        self.dtype_hint = a.dtype_hint


def diag(a, k: int = 0):
    if isinstance(a, Tensor):   # a strided view, like every other view on this path
        d = DiagOp(a, k)
        off = k * a.strides[-1] if k >= 0 else -k * a.strides[-2]
        return Tensor(a.data_ptr + off * A.DTYPE_BYTES[a.dtype], a.dtype, d.shape, tuple(a.strides[:-2]) + (a.strides[-2] + a.strides[-1],), a._keep)
    return DiagOp(a, k)


def isclose(a, b, rtol: float = 1e-5, atol: float = 1e-8):
    """isclose(a, b, rtol, atol) (operators/isclose.h:40-130): int(|a - b| <= atol + rtol * |b|), tolerances in the
    operands' inner type."""
    a = _wrap(a, b if isinstance(b, Op) else None)
    b = _wrap(b, a)
    inner = A.F64 if b.dtype_hint == A.F64 else A.F32
    return as_type(abs(a - b) <= Const(float(atol), inner) + Const(float(rtol), inner) * abs(b), A.I32)


def pow(a, b): return BinOp(A.OP_POW, a, b)  # noqa: A001
def fmod(a, b): return BinOp(A.OP_MOD, a, b)
def atan2(a, b): return BinOp(A.OP_ATAN2, a, b)
def maximum(a, b): return BinOp(A.OP_MAX, a, b)   # MatX: max(a, b) with two operands
def minimum(a, b): return BinOp(A.OP_MIN, a, b)
def as_type(a, dtype: int): return UnOp(A.OP_CAST, a, dtype)


# --------------------------------------------------------------------------------------------------------
# reductions (lazy "transform ops", reference: operators/sum.h etc.)
# --------------------------------------------------------------------------------------------------------
RED_ARGMINMAX = -2   # host-side tag only: lowered to mxb_argminmax, not an mxb_reduce op id


class ReduceExpr:
    def __init__(self, op: int, a, dims: Optional[Sequence[int]], ddof: int = 1):
        a = _wrap(a, None)
        self.op, self.a, self.ddof = op, a, ddof
        r = len(a.shape)
        if dims is None:
            dims = list(range(r))
        dims = [int(d) for d in dims]
        if len(set(dims)) != len(dims) or [d for d in dims if d < 0 or d >= r]:
            raise ValueError("reduction dims must be distinct and inside the operand rank")
        self.dims = dims
        # getPermuteDims (core/utils.h:96-127): batch dims keep their order, reduced dims go last in the listed order
        self.perm = [d for d in range(r) if d not in dims] + dims
        self.out_shape = tuple(a.shape[d] for d in range(r) if d not in dims)


def sum(a, dims=None): return ReduceExpr(A.RED_SUM, a, dims)        # noqa: A001
def mean(a, dims=None): return ReduceExpr(A.RED_MEAN, a, dims)
def var(a, dims=None, ddof: int = 1): return ReduceExpr(A.RED_VAR, a, dims, ddof)
def stdd(a, dims=None, ddof: int = 1): return ReduceExpr(A.RED_STDD, a, dims, ddof)
def max(a, dims=None): return ReduceExpr(A.RED_MAX, a, dims)        # noqa: A001
def min(a, dims=None): return ReduceExpr(A.RED_MIN, a, dims)        # noqa: A001
def argmax(a, dims=None): return ReduceExpr(A.RED_ARGMAX, a, dims)
def argmin(a, dims=None): return ReduceExpr(A.RED_ARGMIN, a, dims)
def argminmax(a, dims=None):
    """`(mtie(minv, mini, maxv, maxi) = argminmax(x, dims))` (operators/argminmax.h): both extrema, one read of x."""
    return ReduceExpr(RED_ARGMINMAX, a, dims)


def any(a, dims=None): return ReduceExpr(A.RED_ANY, a, dims)        # noqa: A001
def all(a, dims=None): return ReduceExpr(A.RED_ALL, a, dims)        # noqa: A001
def prod(a, dims=None): return ReduceExpr(A.RED_PROD, a, dims)


def trace(a): return ReduceExpr(A.RED_SUM, diag(a), None)          # trace_impl, transforms/reduce.h:1505-1511


def allclose(dest: "Tensor", a, b, rtol: float, atol: float, ex: "CudaExecutor") -> None:
    """allclose(dest, in1, in2, rtol, atol, exec) (transforms/reduce.h:1321-1331): runs immediately, rank-0 int output."""
    if len(dest.shape) != 0:
        raise TypeError("allclose output must be rank 0")
    dest.set(all(isclose(a, b, rtol, atol))).run(ex)


class SoftmaxExpr(ReduceExpr):
    """softmax(a, dims) (operators/softmax.h:40-140 -> softmax_impl, transforms/reduce.h:362-445): same rank as the operand."""

    def __init__(self, a, dims):
        super().__init__(-1, a, dims)
        self.out_shape = tuple(self.a.shape)


def softmax(a, dims=None): return SoftmaxExpr(a, dims)


class CumsumExpr:
    """cumsum(a) (operators/cumsum.h -> cumsum_impl, transforms/cub.h:2367-2395): inclusive prefix sum along the LAST
    dim, same rank and sizes as the operand."""

    def __init__(self, a):
        self.a = _wrap(a, None)
        if len(self.a.shape) < 1:
            raise ValueError("cumsum needs an operand of rank >= 1")
        self.out_shape = tuple(self.a.shape)


def cumsum(a): return CumsumExpr(a)


class _Select:
    """Selection functors of find / find_idx (transforms/cub.h:2521-2588): LT{c}(x) = x < c, and so on."""
    op = -1

    def __init__(self, c):
        self.c = c


class LT(_Select): op = A.SEL_LT        # noqa: E701
class GT(_Select): op = A.SEL_GT        # noqa: E701
class EQ(_Select): op = A.SEL_EQ        # noqa: E701
class NEQ(_Select): op = A.SEL_NEQ      # noqa: E701
class LTE(_Select): op = A.SEL_LTE      # noqa: E701
class GTE(_Select): op = A.SEL_GTE      # noqa: E701


class FindExpr:
    """find(a, sel) / find_idx(a, sel) (operators/find.h:40-110, find_idx.h:40-110 -> find_impl / find_idx_impl,
    transforms/cub.h:2609-2625,2705-2721): `(mtie(out, num_found) = find(a, GT{0.5})).run(exec)`; `out` is rank 1,
    `num_found` a rank-0 int tensor; elements are visited in flat row-major order."""

    def __init__(self, a, sel: _Select, want_indices: bool):
        if not isinstance(sel, _Select):
            raise TypeError("find / find_idx take one of LT, GT, EQ, NEQ, LTE, GTE (a callable cannot cross the C ABI)")
        self.a = _wrap(a, None)
        self.sel = sel
        self.want_indices = want_indices


def find(a, sel): return FindExpr(a, sel, False)
def find_idx(a, sel): return FindExpr(a, sel, True)


SORT_DIR_ASC, SORT_DIR_DESC = 0, 1    # SortDirection_t (transforms/cub.h:75-78)


class SortExpr:
    """sort(a, dir) (operators/sort.h -> sort_impl, transforms/cub.h:2145-2190): every row of the last dim sorted, same
    shape as the operand."""

    def __init__(self, a, direction: int = SORT_DIR_ASC):
        self.a = _wrap(a, None)
        if len(self.a.shape) < 1:
            raise ValueError("sort needs an operand of rank >= 1")
        self.direction = direction
        self.out_shape = tuple(self.a.shape)


class HistExpr:
    """hist(a, lower, upper, num_levels) (operators/hist.h:40-125 -> hist_impl, transforms/cub.h:2464-2503): counts of
    every row of the last dim in num_levels - 1 even-width bins over [lower, upper); int output."""

    def __init__(self, a, lower, upper, num_levels: int):
        self.a = _wrap(a, None)
        if len(self.a.shape) < 1:
            raise ValueError("hist needs an operand of rank >= 1")
        self.lower, self.upper, self.num_levels = lower, upper, int(num_levels)
        self.out_shape = tuple(self.a.shape[:-1]) + (self.num_levels - 1,)


class UniqueExpr:
    """unique(a) (operators/unique.h -> unique_impl, transforms/cub.h:2796-2842): `(mtie(out, num_found) = unique(a)).run(exec)`,
    rank-1 operand, the distinct values in ascending order."""

    def __init__(self, a):
        self.a = _wrap(a, None)
        if len(self.a.shape) != 1:
            raise ValueError("unique takes a rank-1 operand")


def sort(a, direction: int = SORT_DIR_ASC): return SortExpr(a, direction)
def hist(a, lower, upper, num_levels: int): return HistExpr(a, lower, upper, num_levels)
def unique(a): return UniqueExpr(a)


class mtie:
    """mtie(values, indices) — multi-output LHS (core/tie.h:44-117)."""

    def __init__(self, *outs: Tensor):
        self.outs = outs

    def set(self, rhs) -> "Set":
        return Set(self, rhs)


# --------------------------------------------------------------------------------------------------------
# lowering to the C ABI
# --------------------------------------------------------------------------------------------------------
class _Lowering:
    def __init__(self, rank: int, size: Sequence[int]):
        self.e = A.Expr()
        self.e.rank = rank
        for d, s in enumerate(size):
            self.e.size[d] = s
        self.memo = {}
        self.keep = []

    def _node(self, opcode, s0, s1=-1, aux=0) -> int:
        i = self.e.n_nodes
        if i >= A.MXB_MAX_NODES:
            raise A.MatxB200Error(A.ERR_NOT_SUPPORTED, "expression has more than %d nodes" % A.MXB_MAX_NODES)
        n = self.e.nodes[i]
        n.opcode, n.src[0], n.src[1], n.aux = opcode, s0, s1, aux
        self.e.n_nodes = i + 1
        return i

    def lower(self, node: Op, axes: Sequence[int]) -> int:
        """axes[i] = dim of the root index space that dim i of `node` walks."""
        key = (id(node), tuple(axes))
        if key in self.memo:
            return self.memo[key]
        if isinstance(node, Tensor):
            k = self.e.n_leaves
            if k >= A.MXB_MAX_LEAVES:
                raise A.MatxB200Error(A.ERR_NOT_SUPPORTED, "expression has more than %d tensor operands" % A.MXB_MAX_LEAVES)
            lf = self.e.leaves[k]
            lf.data, lf.dtype = node.data_ptr, node.dtype
            for d in range(A.MXB_MAX_RANK):
                lf.stride[d] = 0
            for i, ax in enumerate(axes):
                lf.stride[ax] += node.strides[i]
            self.e.n_leaves = k + 1
            self.keep.append(node)
            r = self._node(A.OP_LEAF, k)
        elif isinstance(node, Const):
            k = self.e.n_consts
            if k >= A.MXB_MAX_CONSTS:
                raise A.MatxB200Error(A.ERR_NOT_SUPPORTED, "expression has more than %d constants" % A.MXB_MAX_CONSTS)
            c = self.e.consts[k]
            v = node.value
            c.re, c.im, c.dtype = (v.real, v.imag, node.dtype) if isinstance(v, complex) else (float(v), 0.0, node.dtype)
            self.e.n_consts = k + 1
            r = self._node(A.OP_CONST, k)
        elif isinstance(node, BinOp):
            ra, rb = len(node.a.shape), len(node.b.shape)
            n = len(axes)
            a = self.lower(node.a, axes[n - ra:])
            b = self.lower(node.b, axes[n - rb:])
            r = self._node(node.opcode, a, b)
        elif isinstance(node, UnOp):
            r = self._node(node.opcode, self.lower(node.a, axes), -1, node.aux)
        elif isinstance(node, PermuteOp):
            child_axes = [0] * len(axes)
            for i, d in enumerate(node.dims):
                child_axes[d] = axes[i]
            r = self.lower(node.a, child_axes)
        elif isinstance(node, DiagOp):
            r = self.lower(node.a, list(axes[:-1]) + [axes[-1], axes[-1]])   # rows and columns walk the same root dim
        elif isinstance(node, CloneOp):
            child_axes = [axes[i] for i, c in enumerate(node.cdims) if c == matxKeepDim]
            r = self.lower(node.a, child_axes)
        else:
            raise TypeError("cannot lower %r" % (node,))
        self.memo[key] = r
        return r


def _out_desc(t: Tensor) -> A.Out:
    o = A.Out()
    o.data, o.dtype, o.rank = t.data_ptr, t.dtype, len(t.shape)
    for d, (s, st) in enumerate(zip(t.shape, t.strides)):
        o.size[d], o.stride[d] = s, st
    return o


def lower_elementwise(rhs: Op) -> A.Expr:
    L = _Lowering(len(rhs.shape), rhs.shape)
    L.e.root = L.lower(rhs, list(range(len(rhs.shape))))
    L.e._keep = L.keep
    return L.e


def lower_reduce(r: ReduceExpr) -> A.Expr:
    a = r.a
    shape = tuple(a.shape[d] for d in r.perm)
    L = _Lowering(len(shape), shape)
    # dim i of the permuted view is dim perm[i] of the operand: operand dim perm[i] walks root dim i
    axes = [0] * len(shape)
    for i, d in enumerate(r.perm):
        axes[d] = i
    L.e.root = L.lower(a, axes)
    L.e._keep = L.keep
    return L.e


class Set:
    """`(lhs = rhs)`; `.run(exec)` dispatches like BaseOp::run (operators/base_operator.h:181-285)."""

    def __init__(self, lhs, rhs):
        self.lhs = lhs
        if isinstance(rhs, UniqueExpr):
            if not isinstance(lhs, mtie) or len(lhs.outs) != 2 or len(lhs.outs[0].shape) != 1 or len(lhs.outs[1].shape) != 0:
                raise TypeError("unique needs mtie(out, num_found) with a rank-1 output and a rank-0 count")
            self.rhs = rhs
            return
        if isinstance(rhs, (SortExpr, HistExpr)):
            if isinstance(lhs, mtie):
                raise TypeError("sort / hist have one output")
            if tuple(lhs.shape) != tuple(rhs.out_shape):
                raise A.MatxB200Error(A.ERR_SIZE, "lhs shape %s does not match rhs shape %s" % (lhs.shape, rhs.out_shape))
            self.rhs = rhs
            return
        if isinstance(rhs, FindExpr):
            if not isinstance(lhs, mtie) or len(lhs.outs) != 2:
                raise TypeError("find / find_idx need mtie(out, num_found) on the left-hand side")
            if len(lhs.outs[0].shape) != 1:
                raise TypeError("find output must be rank 1")
            if len(lhs.outs[1].shape) != 0:
                raise TypeError("Num found output tensor rank must be 0")   # the reference's static_assert
            self.rhs = rhs
            return
        self.rhs = rhs if isinstance(rhs, (ReduceExpr, CumsumExpr)) else _wrap(rhs, lhs if isinstance(lhs, Op) else None)
        if isinstance(lhs, mtie):
            if not isinstance(self.rhs, ReduceExpr) or self.rhs.op not in (A.RED_ARGMAX, A.RED_ARGMIN, RED_ARGMINMAX):
                raise TypeError("mtie(...) takes argmax / argmin / argminmax on the right-hand side")
            if len(lhs.outs) != (4 if self.rhs.op == RED_ARGMINMAX else 2):
                raise TypeError("mtie(values, indices) needs two outputs (argminmax: min value, min index, max value, max index)")
        shape = self.rhs.out_shape if isinstance(self.rhs, (ReduceExpr, CumsumExpr)) else self.rhs.shape
        outs = lhs.outs if isinstance(lhs, mtie) else (lhs,)
        for o in outs:
            if tuple(o.shape) != tuple(shape) and not (len(shape) == 0 and len(o.shape) == 0):
                if not isinstance(self.rhs, (ReduceExpr, CumsumExpr)) and len(shape) <= len(o.shape) and tuple(o.shape[len(o.shape) - len(shape):]) == tuple(shape):
                    continue  # lower-rank rhs broadcasts into the lhs
                raise A.MatxB200Error(A.ERR_SIZE, "lhs shape %s does not match rhs shape %s" % (o.shape, shape))  # matxInvalidSize

    def run(self, ex: "CudaExecutor") -> None:
        A.sync_env()
        if isinstance(self.rhs, UniqueExpr):
            e = lower_elementwise(self.rhs.a)
            out, cnt = _out_desc(self.lhs.outs[0]), _out_desc(self.lhs.outs[1])
            A.check(A.lib.mxb_unique(ex.handle, C.byref(e), C.byref(out), C.byref(cnt)))
        elif isinstance(self.rhs, SortExpr):
            e = lower_elementwise(self.rhs.a)
            out = _out_desc(self.lhs)
            A.check(A.lib.mxb_sort(ex.handle, C.byref(e), C.byref(out), 1 if self.rhs.direction == SORT_DIR_DESC else 0))
        elif isinstance(self.rhs, HistExpr):
            e = lower_elementwise(self.rhs.a)
            out = _out_desc(self.lhs)
            A.check(A.lib.mxb_hist(ex.handle, C.byref(e), float(self.rhs.lower), float(self.rhs.upper), C.byref(out)))
        elif isinstance(self.rhs, FindExpr):
            r = self.rhs
            e = lower_elementwise(r.a)
            out, cnt = _out_desc(self.lhs.outs[0]), _out_desc(self.lhs.outs[1])
            A.check(A.lib.mxb_find(ex.handle, C.byref(e), r.sel.op, float(r.sel.c), C.byref(out), C.byref(cnt), 1 if r.want_indices else 0))
        elif isinstance(self.rhs, CumsumExpr):
            e = lower_elementwise(self.rhs.a)
            out = _out_desc(self.lhs)
            A.check(A.lib.mxb_cumsum(ex.handle, C.byref(e), C.byref(out)))
        elif isinstance(self.rhs, SoftmaxExpr):
            r = self.rhs
            e = lower_reduce(r)
            o = _out_desc(self.lhs)
            out = A.Out()
            out.data, out.dtype, out.rank = o.data, o.dtype, o.rank
            for i, d in enumerate(r.perm):   # the output is walked in the same permuted order as the operand
                out.size[i], out.stride[i] = o.size[d], o.stride[d]
            A.check(A.lib.mxb_softmax(ex.handle, C.byref(e), len(r.dims), C.byref(out)))
        elif isinstance(self.rhs, ReduceExpr):
            r = self.rhs
            e = lower_reduce(r)
            if r.op == RED_ARGMINMAX:
                if not isinstance(self.lhs, mtie):
                    raise TypeError("argminmax needs mtie(minv, mini, maxv, maxi) on the left-hand side")
                o = [_out_desc(t) for t in self.lhs.outs]
                A.check(A.lib.mxb_argminmax(ex.handle, C.byref(e), len(r.dims), C.byref(o[0]), C.byref(o[1]), C.byref(o[2]), C.byref(o[3])))
            elif isinstance(self.lhs, mtie):
                out, idx = _out_desc(self.lhs.outs[0]), _out_desc(self.lhs.outs[1])
                A.check(A.lib.mxb_reduce(ex.handle, r.op, C.byref(e), len(r.dims), C.byref(out), C.byref(idx), r.ddof))
            else:
                if r.op in (A.RED_ARGMAX, A.RED_ARGMIN):
                    raise TypeError("argmax / argmin need mtie(values, indices) on the left-hand side")
                out = _out_desc(self.lhs)
                A.check(A.lib.mxb_reduce(ex.handle, r.op, C.byref(e), len(r.dims), C.byref(out), None, r.ddof))
        else:
            rhs = self.rhs
            lshape = self.lhs.shape
            if len(rhs.shape) < len(lshape):
                rhs = CloneOp(rhs, list(lshape[:len(lshape) - len(rhs.shape)]) + [matxKeepDim] * len(rhs.shape))
            e = lower_elementwise(rhs)
            out = _out_desc(self.lhs)
            A.check(A.lib.mxb_elementwise(ex.handle, C.byref(e), C.byref(out)))


# --------------------------------------------------------------------------------------------------------
# executor
# --------------------------------------------------------------------------------------------------------
class CudaExecutor:
    """Mirror of matx::cudaExecutor (executors/cuda.h:60-82, cuda_executor_common.h:85-176): a stream plus
    sync() and event timers.  Owns one library handle (grid-combine scratch lives there)."""

    def __init__(self, stream=None):
        if stream is None:
            ptr = 0
        elif isinstance(stream, int):
            ptr = stream
        else:
            ptr = int(stream.cuda_stream)  # torch.cuda.Stream
        self._stream_ptr = ptr
        h = C.c_void_p()
        A.check(A.lib.mxb_create(C.byref(h), C.c_void_p(ptr)))
        self.handle = h
        self._ev = None

    def getStream(self) -> int:
        return self._stream_ptr

    def set_stream(self, stream) -> None:
        """Rebind the executor (and its library handle) to another stream (torch.cuda.Stream, raw pointer or None)."""
        ptr = 0 if stream is None else (stream if isinstance(stream, int) else int(stream.cuda_stream))
        A.check(A.lib.mxb_set_stream(self.handle, C.c_void_p(ptr)))
        self._stream_ptr = ptr

    def sync(self) -> None:
        A.check(A.lib.mxb_sync(self.handle))

    def _torch_stream(self):
        import torch
        return torch.cuda.ExternalStream(self._stream_ptr) if self._stream_ptr else torch.cuda.default_stream()

    def start_timer(self) -> None:
        import torch
        self._ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        self._ev[0].record(self._torch_stream())

    def stop_timer(self) -> None:
        self._ev[1].record(self._torch_stream())

    def get_time_ms(self) -> float:
        self._ev[1].synchronize()
        return self._ev[0].elapsed_time(self._ev[1])

    def last_kernel(self) -> str:
        return (A.lib.mxb_last_kernel(self.handle) or b"").decode()

    def launch_count(self) -> int:
        return int(A.lib.mxb_launch_count(self.handle))

    def close(self) -> None:
        if self.handle:
            A.lib.mxb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
