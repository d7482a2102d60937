"""A/B of the full-tensor argmax / var deal shapes on one box (development tool, run under gpurun)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

for n in (1 << 30,):
    x = torch.rand(n, device="cuda")
    tx = mx.make_tensor(x)
    o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    o2, oi2 = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    for name, fn in (("argmax", lambda ex: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx)).run(ex)),
                     ("var", lambda ex: mx.make_tensor(o).set(mx.var(tx, None, 1)).run(ex)),
                     ("argminmax", lambda ex: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi), mx.make_tensor(o2), mx.make_tensor(oi2)).set(mx.argminmax(tx)).run(ex)),
                     ("max", lambda ex: mx.make_tensor(o).set(mx.max(tx)).run(ex))):
        for env in ({}, {"MXB_TUNE_MINB": "4"}, {"MXB_TUNE_V": "8", "MXB_TUNE_U": "4"}, {"MXB_TUNE_V": "8", "MXB_TUNE_U": "4", "MXB_TUNE_MINB": "4"},
                    {"MXB_TUNE_V": "8", "MXB_TUNE_U": "2", "MXB_TUNE_MINB": "4"}):
            if name in ("var", "max") and env:
                continue
            old = {k: os.environ.get(k) for k in env}
            os.environ.update(env)
            ex = mx.CudaExecutor()
            ms, best = bc._time(ex, lambda: fn(ex), iters=10, warm=3)
            print(json.dumps({"n": n, "op": name, "env": env, "ms": round(ms, 4), "GBps": round(n * 4 / ms / 1e6, 1), "kernel": ex.last_kernel()}), flush=True)
            for k, v in old.items():
                os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    del x, tx
    torch.cuda.empty_cache()
