"""ctypes mirror of include/matx_b200.h and the loader of libmatx_b200.so.

The library is the product: if it is missing or fails to load, the first use of `lib` raises (and so does
`matx_b200.ops`'s CudaExecutor) — there is no Python / torch fallback for any compute entry point.  The constants and
struct mirrors above the loader are importable without the library (bench.py's reference arm uses them and must not
map libmatx_b200.so into its process).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmatx_b200.so")

MXB_MAX_RANK = 8
MXB_MAX_LEAVES = 12
MXB_MAX_NODES = 96
MXB_MAX_CONSTS = 24
MXB_PARTIAL_BYTES = 32

# mxb_status_t
OK, ERR_INVALID, ERR_NOT_SUPPORTED, ERR_CUDA, ERR_NO_DEVICE, ERR_JIT, ERR_SIZE = range(7)
STATUS_NAMES = ["MXB_OK", "MXB_ERR_INVALID", "MXB_ERR_NOT_SUPPORTED", "MXB_ERR_CUDA", "MXB_ERR_NO_DEVICE", "MXB_ERR_JIT",
                "MXB_ERR_SIZE"]

# mxb_dtype_t
F32, F64, BF16, F16, C64, I32, I64, U8 = range(8)
DTYPE_NAMES = ["f32", "f64", "bf16", "f16", "c64", "i32", "i64", "u8"]
DTYPE_BYTES = [4, 8, 2, 2, 8, 4, 8, 1]

# mxb_reduce_op_t
RED_SUM, RED_MEAN, RED_VAR, RED_STDD, RED_MAX, RED_MIN, RED_ARGMAX, RED_ARGMIN, RED_ANY, RED_ALL, RED_PROD = range(11)
SEL_LT, SEL_GT, SEL_EQ, SEL_NEQ, SEL_LTE, SEL_GTE = range(6)   # mxb_select_op_t

# mxb_opcode_t
OP_LEAF, OP_CONST = 0, 1
(OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_MOD, OP_POW, OP_MAX, OP_MIN, OP_LT, OP_GT, OP_LE, OP_GE, OP_EQ, OP_NE, OP_AND, OP_OR,
 OP_ATAN2) = range(10, 27)
(OP_NEG, OP_SQRT, OP_RSQRT, OP_EXP, OP_LOG, OP_LOG2, OP_LOG10, OP_ABS, OP_ABS2, OP_CONJ, OP_REAL, OP_IMAG, OP_SIN, OP_COS,
 OP_TAN, OP_TANH, OP_NORMCDF, OP_NOT, OP_ISNAN, OP_ISINF, OP_FLOOR, OP_CEIL, OP_ROUND, OP_SINH, OP_COSH, OP_ASIN, OP_ACOS,
 OP_ATAN, OP_EXPJ, OP_CSQRT_UNUSED) = range(40, 70)
OP_CAST = 80


class Node(C.Structure):
    _fields_ = [("opcode", C.c_int32), ("src", C.c_int32 * 2), ("aux", C.c_int32)]


class Leaf(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("_pad", C.c_int32), ("stride", C.c_int64 * MXB_MAX_RANK)]


class Const(C.Structure):
    _fields_ = [("re", C.c_double), ("im", C.c_double), ("dtype", C.c_int32), ("_pad", C.c_int32)]


class Expr(C.Structure):
    _fields_ = [("rank", C.c_int32), ("n_nodes", C.c_int32), ("n_leaves", C.c_int32), ("n_consts", C.c_int32),
                ("root", C.c_int32), ("_pad", C.c_int32), ("size", C.c_int64 * MXB_MAX_RANK),
                ("nodes", Node * MXB_MAX_NODES), ("leaves", Leaf * MXB_MAX_LEAVES), ("consts", Const * MXB_MAX_CONSTS)]


class Out(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("rank", C.c_int32), ("size", C.c_int64 * MXB_MAX_RANK),
                ("stride", C.c_int64 * MXB_MAX_RANK)]


# every symbol include/matx_b200.h declares (tests check the library exports all of them)
EXPORTED = [
    "mxb_create", "mxb_destroy", "mxb_set_stream", "mxb_sync", "mxb_elementwise", "mxb_reduce", "mxb_reduce_partial",
    "mxb_reduce_finalize", "mxb_version", "mxb_last_error", "mxb_device_count", "mxb_last_kernel", "mxb_launch_count",
    "mxb_is_aot", "mxb_reload_env", "mxb_reduce_partial_push", "mxb_exchange_finalize", "mxb_exchange_alloc", "mxb_exchange_open", "mxb_exchange_close",
    "mxb_exchange_free", "mxb_exchange_check", "mxb_softmax", "mxb_cumsum", "mxb_find", "mxb_hist", "mxb_sort", "mxb_unique", "mxb_argminmax",
]


MXB_MAX_PEERS = 8
MXB_MAX_ITEMS = 8


class Peers(C.Structure):
    _fields_ = [("rec", C.c_void_p * MXB_MAX_PEERS), ("flag", C.c_void_p * MXB_MAX_PEERS), ("epoch", C.c_void_p),
                ("world", C.c_int32), ("rank", C.c_int32)]


class FoldItem(C.Structure):
    _fields_ = [("reduce_op", C.c_int32), ("value_dtype", C.c_int32), ("out", C.c_void_p), ("idx_out", C.c_void_p),
                ("ddof", C.c_int32), ("_pad", C.c_int32)]


def exchange_rec_bytes(world: int) -> int:
    return 2 * world * MXB_MAX_ITEMS * MXB_PARTIAL_BYTES


class MatxB200Error(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__("%s: %s" % (STATUS_NAMES[status] if 0 <= status < len(STATUS_NAMES) else status, message))
        self.status = status


def _load() -> C.CDLL:
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "matx_b200: %s is missing. Build it with `python -m matx_b200.build` (needs nvcc). "
            "There is no CPU or PyTorch fallback for this path." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64 = C.c_void_p, C.c_int32, C.c_int64
    lib.mxb_create.argtypes = [C.POINTER(vp), vp]
    lib.mxb_destroy.argtypes = [vp]
    lib.mxb_set_stream.argtypes = [vp, vp]
    lib.mxb_sync.argtypes = [vp]
    lib.mxb_elementwise.argtypes = [vp, C.POINTER(Expr), C.POINTER(Out)]
    lib.mxb_reduce.argtypes = [vp, i32, C.POINTER(Expr), i32, C.POINTER(Out), C.POINTER(Out), i32]
    lib.mxb_softmax.argtypes = [vp, C.POINTER(Expr), i32, C.POINTER(Out)]
    lib.mxb_cumsum.argtypes = [vp, C.POINTER(Expr), C.POINTER(Out)]
    lib.mxb_find.argtypes = [vp, C.POINTER(Expr), i32, C.c_double, C.POINTER(Out), C.POINTER(Out), i32]
    lib.mxb_argminmax.argtypes = [vp, C.POINTER(Expr), i32, C.POINTER(Out), C.POINTER(Out), C.POINTER(Out), C.POINTER(Out)]
    lib.mxb_hist.argtypes = [vp, C.POINTER(Expr), C.c_double, C.c_double, C.POINTER(Out)]
    lib.mxb_sort.argtypes = [vp, C.POINTER(Expr), C.POINTER(Out), i32]
    lib.mxb_unique.argtypes = [vp, C.POINTER(Expr), C.POINTER(Out), C.POINTER(Out)]
    lib.mxb_reduce_partial.argtypes = [vp, i32, C.POINTER(Expr), i64, vp]
    lib.mxb_reduce_finalize.argtypes = [vp, i32, i32, vp, i32, i64, i64, i32, C.POINTER(Out), C.POINTER(Out)]
    lib.mxb_reduce_partial_push.argtypes = [vp, i32, C.POINTER(Expr), i64, C.POINTER(Peers), i32, i32]
    lib.mxb_exchange_finalize.argtypes = [vp, C.POINTER(Peers), C.POINTER(FoldItem), i32, i64]
    lib.mxb_exchange_check.argtypes = [vp, C.POINTER(Peers)]
    lib.mxb_exchange_alloc.argtypes = [vp, C.c_size_t, C.POINTER(vp), C.c_char_p]
    lib.mxb_exchange_open.argtypes = [vp, C.c_char_p, C.POINTER(vp)]
    lib.mxb_exchange_close.argtypes = [vp, vp]
    lib.mxb_exchange_free.argtypes = [vp, vp]
    lib.mxb_reload_env.restype = None
    lib.mxb_last_error.restype = C.c_char_p
    lib.mxb_last_kernel.argtypes = [vp]
    lib.mxb_last_kernel.restype = C.c_char_p
    lib.mxb_launch_count.argtypes = [vp]
    lib.mxb_launch_count.restype = i64
    lib.mxb_is_aot.argtypes = [C.POINTER(Expr), i32]
    lib.mxb_debug_codegen.argtypes = [C.POINTER(Expr), C.c_char_p, C.c_size_t]
    lib.mxb_debug_compile.argtypes = [C.POINTER(Expr), i32, i32, i32, i32, i32, C.c_char_p, C.c_size_t]
    for name in EXPORTED:
        fn = getattr(lib, name)
        if fn.restype is C.c_int:
            fn.restype = C.c_int
    return lib


_env_seen = None


def sync_env() -> None:
    """The library reads its MXB_* knobs once per thread; tests flip them with monkeypatch between statements, so the Python
    mirror tells the library to re-read whenever the MXB_* part of os.environ has changed since the last statement."""
    global _env_seen
    cur = tuple(sorted((k, v) for k, v in os.environ.items() if k.startswith("MXB_")))
    if cur != _env_seen:
        if _env_seen is not None or cur:
            load().mxb_reload_env()
        _env_seen = cur


_lib = None


def load() -> C.CDLL:
    """dlopen libmatx_b200.so once; raises ImportError when it is missing."""
    global _lib
    if _lib is None:
        _lib = _load()
    return _lib


def __getattr__(name: str):
    if name == "lib":   # `A.lib` loads on first use
        return load()
    raise AttributeError("module %r has no attribute %r" % (__name__, name))


def check(status: int) -> None:
    if status != OK:
        raise MatxB200Error(status, (load().mxb_last_error() or b"").decode())
