"""cumsum timing: one row of 2^28 fp32 and 4096 x 8192, flat exchange vs the pipelined one (development tool, run under gpurun)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

PEAK = 6456.8
envs = [{}] + [json.loads(a) for a in sys.argv[1:]]
for shape in ((1 << 28,), (4, 1 << 26), (3, 5000001)):
    x = torch.rand(shape, device="cuda")
    out = torch.empty_like(x)
    want = torch.cumsum(x.double(), -1)
    for env in envs:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        ex = mx.CudaExecutor()
        fn = lambda: mx.make_tensor(out).set(mx.cumsum(mx.make_tensor(x))).run(ex)  # noqa: E731
        ms, best = bc._time(ex, fn, iters=10, warm=3)
        err = float(((out.double() - want).abs() / want.abs().clamp_min(1.0)).max().item())
        nbytes = x.numel() * 8
        print(json.dumps({"shape": list(shape), "env": env, "ms": round(ms, 4), "best": round(best, 4), "rel_err": err,
                          "GBps": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / PEAK, 3), "kernel": ex.last_kernel()}), flush=True)
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
