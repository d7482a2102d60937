"""Run in a subprocess with MXB_PLAN_ONLY=1 (tests/test_dispatch_plan.py): states the BASELINE.json configurations and a
few neighbours at FULL SIZE on untouched host buffers, lets the library plan each statement (no CUDA call, nothing
launched, nothing computed) and prints {case: "kernel key|grid=..|block=..|smem=.."} as one JSON line."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
assert os.environ.get("MXB_PLAN_ONLY") == "1"
from matx_b200 import _abi as A  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402
from tests.oracle_harness import np_tensor  # noqa: E402

ex = mx.CudaExecutor()
res = {}


def empty(shape, dt):
    """Untouched virtual memory (no page is ever read or written), 256-byte aligned like a device allocation."""
    n = int(np.prod(shape)) if shape != () else 1
    item = np.dtype(dt).itemsize
    raw = np.empty(n * item + 256, np.uint8)
    off = (-raw.ctypes.data) % 256
    return raw[off:off + n * item].view(dt).reshape(shape)


def plan(name, lhs, rhs):
    (lhs.set(rhs) if not isinstance(lhs, tuple) else mx.mtie(*lhs).set(rhs)).run(ex)
    res[name] = ex.last_kernel()


f32, c64 = np.float32, np.complex64
o0 = np_tensor(empty((), f32))
i0 = np_tensor(empty((), np.int64))
x30 = np_tensor(empty(1 << 30, f32))
plan("c2.sum", o0, mx.sum(x30))
plan("c2.argmax", (o0, i0), mx.argmax(x30))
plan("full.var", o0, mx.var(x30, None, 1))
a = np_tensor(empty((16384, 4096), f32))
b1, c1 = np_tensor(empty((16384, 4096), f32)), np_tensor(empty((16384, 4096), f32))
plan("c1.fma_sum", np_tensor(empty(16384, f32)), mx.sum(a * b1 + c1, [1]))
x3 = np_tensor(empty((65536, 8192), c64))
plan("c3.mean", np_tensor(empty(65536, c64)), mx.mean(x3, [1]))
plan("c3.var", np_tensor(empty(65536, f32)), mx.var(x3, [1], 1))
plan("c3.argmax_abs2", (np_tensor(empty(65536, f32)), np_tensor(empty(65536, np.int64))), mx.argmax(mx.abs2(x3), [1]))
t5 = np_tensor(empty((1024, 1024, 1024), np.uint16), A.BF16)
plan("c5.permuted_sum", np_tensor(empty((1024, 1024), np.uint16), A.BF16), mx.sum(mx.permute(t5, [2, 0, 1]), [2]))
m = np_tensor(empty((4096, 65536), f32))
plan("colsum.4096x65536", np_tensor(empty(65536, f32)), mx.sum(m, [0]))
plan("colvar.4096x65536", np_tensor(empty(65536, f32)), mx.var(m, [0], 1))
plan("rowvar.4096x65536", np_tensor(empty(4096, f32)), mx.var(m, [1], 1))
tall = np_tensor(empty((3000, 1000), f32))
plan("colsum.tall", np_tensor(empty(1000, f32)), mx.sum(tall, [0]))
short = np_tensor(empty((1 << 20, 64), f32))
plan("rowvar.short64", np_tensor(empty(1 << 20, f32)), mx.var(short, [1], 1))
plan("rowvar.1024", np_tensor(empty(1 << 16, f32)), mx.var(np_tensor(empty((1 << 16, 1024), f32)), [1], 1))
plan("rowsum.short64", np_tensor(empty(1 << 20, f32)), mx.sum(short, [1]))
v = np_tensor(empty(1 << 28, f32))
plan("ew.vector_add", np_tensor(empty(1 << 28, f32)), v + np_tensor(empty(1 << 28, f32)))
p4 = np_tensor(empty((1000, 200, 6, 300), f32))
plan("ew_tr.permute", np_tensor(empty((300, 1000, 6, 200), f32)), p4.Permute([3, 0, 2, 1]))
plan("scan.rows", np_tensor(empty((16384, 4096), f32)), mx.cumsum(a))
plan("find.values", (np_tensor(empty(1 << 28, f32)), np_tensor(empty((), np.int32))), mx.find(v, mx.GT(0.5)))
plan("find.strided_idx", (np_tensor(empty(1 << 20, np.int32)), np_tensor(empty((), np.int32))),
     mx.find_idx(np_tensor(empty((2048, 1024), f32)[:, ::2]), mx.LT(0.5)))
# eligibility edges of the TMA-tiled reduce_outer
narrow = np_tensor(empty((300, 64, 36), f32))
plan("tma.narrow_rows", np_tensor(empty((36, 300), f32)), mx.max(mx.permute(narrow, [2, 0, 1]), [2]))          # 9 chunks per row
shortr = np_tensor(empty((400, 32, 512), f32))
plan("tma.short_reduce_dim", np_tensor(empty((512, 400), f32)), mx.max(mx.permute(shortr, [2, 0, 1]), [2]))   # R = 32 < 64
m2 = np_tensor(empty((4096, 65536), f32))
plan("tma.fused_expression", np_tensor(empty(65536, f32)), mx.sum(m * m2, [0]))                              # two leaves
odd = np_tensor(empty((4096, 65536 + 4), f32)[:, 1:65533])                                                   # rows start 4 bytes off
plan("tma.unaligned", np_tensor(empty(65532, f32)), mx.sum(odd, [0]))
print(json.dumps(res))
