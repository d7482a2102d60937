"""A compact tour of every kernel family on small shapes, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from matx_b200 import _abi as A
from matx_b200 import dist as mxd
from matx_b200 import ops as mx

ex = mx.CudaExecutor()
rng = np.random.default_rng(0)


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def red(r, dtype=torch.float32, idx=False):
    o = torch.zeros(r.out_shape, dtype=dtype, device="cuda")
    if idx:
        i = torch.zeros(r.out_shape, dtype=torch.int64, device="cuda")
        mx.mtie(mx.make_tensor(o), mx.make_tensor(i)).set(r).run(ex)
    else:
        mx.make_tensor(o).set(r).run(ex)
    ex.sync()
    return ex.last_kernel()


x = mx.make_tensor(dev(rng.random((37, 4100), dtype=np.float32)))          # ragged tails, CTA team
y = mx.make_tensor(dev(rng.random((300, 130), dtype=np.float32)))           # warp team, V=1 pitch
big = mx.make_tensor(dev(rng.random(3_000_001, dtype=np.float32)))          # splits + in-launch grid combine
c = mx.make_tensor(dev((rng.standard_normal((24, 8192)) + 1j * rng.standard_normal((24, 8192))).astype(np.complex64)))
t3 = mx.make_tensor(dev(rng.random((6, 33, 48), dtype=np.float32)))
seen = set()
for f in (mx.sum, mx.max, mx.argmax, mx.argmin, mx.any, mx.all, mx.prod):
    for src, dims in ((x, [1]), (x, [0]), (y, [1]), (y, [0]), (big, None), (t3, [0, 2]), (t3, [1])):
        seen.add(red(f(src, dims), idx=f in (mx.argmax, mx.argmin)).split("|")[0])
seen.add(red(mx.var(y, [1])).split("|")[0])                                 # var_group (short rows)
seen.add(red(mx.var(x * 2.0, [1])).split("|")[0])                           # var_reg (fused expression)
tall = mx.make_tensor(dev(rng.random((100_000, 24), dtype=np.float32)))   # reduce_outer with splits
for f in (mx.sum, mx.argmax):
    seen.add(red(f(tall, [0]), idx=f is mx.argmax).split("|")[0])
seen.add(red(mx.var(x, [1])).split("|")[0])                                 # var_tma (16 KB rows)
seen.add(red(mx.var(c, [1])).split("|")[0])                                 # var_tma
os.environ["MXB_VAR_SMEM_ONLY"] = "1"
seen.add(red(mx.stdd(c, [1])).split("|")[0])                                # var_smem
os.environ.pop("MXB_VAR_SMEM_ONLY")
seen.add(red(mx.var(big)).split("|")[0])                                    # two-launch variance
seen.add(red(mx.mean(c, [1]), dtype=torch.complex64).split("|")[0])
seen.add(red(mx.argmax(mx.abs2(c), [1]), idx=True).split("|")[0])
o = torch.zeros((37, 4100), device="cuda")
mx.make_tensor(o).set(mx.sqrt(x) * 2.0 + mx.exp(-x)).run(ex)               # elementwise vector path
o2 = torch.zeros((4100, 37), device="cuda")
mx.make_tensor(o2).set(mx.permute(x, [1, 0]) + 1.0).run(ex)                 # elementwise scalar path
ex.sync()
seen.add(ex.last_kernel().split("|")[0])
# fused exchange, 3 simulated ranks
world = 3
bufs = [torch.zeros(mxd.PeerExchange.buffer_bytes(world), dtype=torch.uint8, device="cuda") for _ in range(world)]
pes = [mxd.PeerExchange(ex, world, r, _sim_buffers=bufs) for r in range(world)]
flat = dev(rng.random(90_000, dtype=np.float32))
outs = []
import ctypes as C
plans = []
for r in range(world):
    s, cnt = mxd.slab(90_000, r, world, align=64)
    oo = [torch.zeros((), device="cuda"), torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda"), torch.zeros((), device="cuda")]
    plans.append((pes[r], pes[r].prepare([(A.RED_SUM, oo[0], None), (A.RED_ARGMAX, oo[1], oo[2]), (A.RED_VAR, oo[3], None)], mx.make_tensor(flat[s:s + cnt]), s, 90_000)))
for pe, plan in plans:
    for op, e, off, k in plan["push"]:
        A.check(A.lib.mxb_reduce_partial_push(ex.handle, op, C.byref(e), off, C.byref(pe.peers), k, plan["n"]))
for pe, plan in plans:
    A.check(A.lib.mxb_exchange_finalize(ex.handle, C.byref(pe.peers), plan["fold"], plan["n"], plan["count"]))
ex.sync()
print("kernel families exercised:", sorted(seen), "launches:", ex.launch_count())
