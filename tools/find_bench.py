"""find / find_idx timing on 2^28 fp32 at several selectivities and launch shapes (development tool, run under gpurun).
Prints one JSON line per setting; `torch` column = torch.masked_select-free baseline is not used — the A/B against
cub::DeviceSelect is tests/cpp/dropin_test --bench."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

PEAK = 6456.8
n = 1 << 28
x = torch.rand(n, device="cuda")
tx = mx.make_tensor(x)
out = torch.empty(n, device="cuda")
iout = torch.empty(n, dtype=torch.int32, device="cuda")
nf = torch.zeros((), dtype=torch.int32, device="cuda")
envs = [{}] + [{"MXB_TUNE_SEL_CTAS": str(c)} for c in (1, 2)]
if len(sys.argv) > 1:
    envs = [json.loads(a) for a in sys.argv[1:]]
for env in envs:
    old = {k: os.environ.get(k) for k in env}
    os.environ.update(env)
    ex = mx.CudaExecutor()
    for thr in (0.99, 0.9, 0.5, 0.0):
        for idx in (False, True):
            o = iout if idx else out
            f = (mx.find_idx if idx else mx.find)(tx, mx.GT(thr))
            fn = lambda: mx.mtie(mx.make_tensor(o), mx.make_tensor(nf)).set(f).run(ex)  # noqa: E731
            ms, best = bc._time(ex, fn, iters=10, warm=3)
            sel = int(nf.item())
            want = int((x > thr).sum().item())
            nbytes = n * 4 + sel * 4 + 4
            print(json.dumps({"thr": thr, "idx": idx, "env": env, "ms": round(ms, 4), "best": round(best, 4), "selected": sel, "count_ok": sel == want,
                              "GBps": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / PEAK, 3), "kernel": ex.last_kernel()}), flush=True)
    for k, v in old.items():
        os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
