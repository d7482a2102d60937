"""One launch of every headline kernel (configs 1-5), for `ncu` captures under gpurun.  Not a benchmark:
numbers printed under a profiler are never reported."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

ex = mx.CudaExecutor()
which = sys.argv[1:] or ["c2", "c1", "c3", "c4", "c5"]
bc._time = lambda ex_, fn, iters=1, warm=0: (fn(), torch.cuda.synchronize(), (1.0, 1.0))[2]  # single launch each
if "c2" in which:
    x = torch.rand(1 << 30, device="cuda")
    tx = mx.make_tensor(x)
    o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    mx.make_tensor(o).set(mx.sum(tx)).run(ex)
    mx.make_tensor(o).set(mx.max(tx)).run(ex)
    mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx)).run(ex)
    ex.sync()
    del x, tx
    torch.cuda.empty_cache()
for name, f in (("c1", bc.run_c1), ("c3", bc.run_c3), ("c4", bc.run_c4), ("c5", bc.run_c5)):
    if name in which:
        f(ex, 6456.8)
        torch.cuda.empty_cache()
if "perm" in which:   # the reference's permute benchmark shape + a plain transpose, through the transposing family
    x = torch.randn(1000, 200, 6, 300, device="cuda")
    y = torch.empty(300, 1000, 6, 200, device="cuda")
    mx.make_tensor(y).set(mx.make_tensor(x).Permute([3, 0, 2, 1])).run(ex)
    a = torch.randn(8192, 8192, device="cuda")
    t = torch.empty(8192, 8192, device="cuda")
    mx.make_tensor(t).set(mx.make_tensor(a).Permute([1, 0])).run(ex)
    ex.sync()
if "scan" in which:   # cumsum: many rows (row walker), one long row (tile exchange), short rows (warp per row)
    for shape in ((16384, 4096), (1 << 28,), (1 << 18, 256)):
        x = torch.rand(*shape, device="cuda")
        y = torch.empty_like(x)
        mx.make_tensor(y).set(mx.cumsum(mx.make_tensor(x))).run(ex)
        ex.sync()
        del x, y
        torch.cuda.empty_cache()
print("profiled", which)
