"""Single-row cumsum (2^28 fp32, the tile-exchange mode): grid size, prefetch position and tile size.  Development
tool, run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import bench_configs as bc, ops as mx

ex = mx.CudaExecutor()
n = 1 << 28
x = (torch.rand(n, device="cuda") * 2 - 1).round() * 3
y = torch.empty_like(x)
tx, ty = mx.make_tensor(x), mx.make_tensor(y)
want_last = float(x.double().sum())
for env in [{}, {"MXB_SCAN_FLAGS": 2}, {"MXB_SCAN_GRID_PER_SM": 2}, {"MXB_SCAN_ONE_TILE_PER_CTA": 1}, {"MXB_TUNE_U": 8}, {"MXB_TUNE_U": 2}]:
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        ms, best = bc._time(ex, lambda: ty.set(mx.cumsum(tx)).run(ex), iters=6, warm=2)
        ok = float(y[-1]) == want_last and bool(torch.equal(y[1:] - y[:-1], x[1:]))
        print(json.dumps({"env": env, "ms": round(ms, 4), "best": round(best, 4), "GBps": round(2 * n * 4 / ms / 1e6, 1), "exact": ok, "kernel": ex.last_kernel()}), flush=True)
    except Exception as exc:
        print(json.dumps({"env": env, "error": str(exc)[:200]}), flush=True)
    for k in env:
        os.environ.pop(k)
