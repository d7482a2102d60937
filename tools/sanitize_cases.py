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
# ---- round 2: argminmax, single-pass find / find_idx / unique (warp tiles, ring in shared memory), sort (bitonic rows and
# the radix pass with its shared-memory staging), hist, cumsum in all three modes ----
mn, mxv = torch.zeros(37, device="cuda"), torch.zeros(37, device="cuda")
imn, imx = torch.zeros(37, dtype=torch.int64, device="cuda"), torch.zeros(37, dtype=torch.int64, device="cuda")
mx.mtie(*(mx.make_tensor(t) for t in (mn, imn, mxv, imx))).set(mx.argminmax(x, [1])).run(ex)
ex.sync()
seen.add(ex.last_kernel().split("|")[0] + ":" + ex.last_kernel().split("|")[2])
nf = torch.zeros((), dtype=torch.int32, device="cuda")
for n_el, thr in ((300_001, 0.99), (300_001, 0.5), (300_001, 0.0), (5000, 0.9)):
    fx = dev(rng.random(n_el, dtype=np.float32))
    fo = torch.zeros(n_el, device="cuda")
    fi = torch.zeros(n_el, dtype=torch.int64, device="cuda")
    mx.mtie(mx.make_tensor(fo), mx.make_tensor(nf)).set(mx.find(mx.make_tensor(fx), mx.GT(thr))).run(ex)
    mx.mtie(mx.make_tensor(fi), mx.make_tensor(nf)).set(mx.find_idx(mx.make_tensor(fx), mx.GT(thr))).run(ex)
    ex.sync()
    seen.add(ex.last_kernel().split("|")[0] + ":" + ex.last_kernel().split("|")[6])
strided = mx.make_tensor(dev(rng.random((1000, 64), dtype=np.float32))[:, ::2])      # scalar walk of the single-pass kernel
so = torch.zeros(32000, device="cuda")
mx.mtie(mx.make_tensor(so), mx.make_tensor(nf)).set(mx.find(strided, mx.LT(0.3))).run(ex)
keys = dev(rng.integers(-500, 500, 200_000).astype(np.float32))
srt, uq = torch.zeros(200_000, device="cuda"), torch.zeros(200_000, device="cuda")
mx.make_tensor(srt).set(mx.sort(mx.make_tensor(keys))).run(ex)
ex.sync()
seen.add(ex.last_kernel().split("|")[0])
mx.mtie(mx.make_tensor(uq), mx.make_tensor(nf)).set(mx.unique(mx.make_tensor(keys))).run(ex)
rows = dev(rng.random((64, 1000), dtype=np.float32))
srows = torch.zeros((64, 1000), device="cuda")
mx.make_tensor(srows).set(mx.sort(mx.make_tensor(rows), mx.SORT_DIR_DESC)).run(ex)
ex.sync()
seen.add(ex.last_kernel().split("|")[0])
k64 = dev(rng.standard_normal(70_000))
s64 = torch.zeros(70_000, dtype=torch.float64, device="cuda")
mx.make_tensor(s64).set(mx.sort(mx.make_tensor(k64))).run(ex)
hb = torch.zeros(16, dtype=torch.int32, device="cuda")
mx.make_tensor(hb).set(mx.hist(mx.make_tensor(keys), -125.0, 125.0, 17)).run(ex)
ex.sync()
seen.add(ex.last_kernel().split("|")[0])
for shape in ((3_000_001,), (300, 130), (24, 8192)):
    cx = dev(rng.random(shape, dtype=np.float32))
    co = torch.zeros(shape, device="cuda")
    mx.make_tensor(co).set(mx.cumsum(mx.make_tensor(cx))).run(ex)
    ex.sync()
    seen.add(ex.last_kernel().split("|")[0] + ":" + ex.last_kernel().split("|")[6])
print("kernel families exercised:", sorted(seen), "launches:", ex.launch_count())
