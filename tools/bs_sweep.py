"""Black-Scholes (config 4) launch-shape sweep: grid as one batch per thread vs a persistent grid of k CTAs per SM, and
the unroll U.  Development tool, run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import bench_configs as bc, ops as mx

ex = mx.CudaExecutor()
n = 1 << 28
S = torch.rand(n, device="cuda") * 90 + 10
K = torch.rand(n, device="cuda") * 90 + 10
V = torch.rand(n, device="cuda") * 0.45 + 0.05
r = torch.rand(n, device="cuda") * 0.09 + 0.01
T = torch.rand(n, device="cuda") * 1.9 + 0.1
out = torch.empty(n, device="cuda")
tK, tS, tV, tr, tT, to = (mx.make_tensor(t) for t in (K, S, V, r, T, out))
expr = bc.black_scholes_expr(tK, tS, tV, tr, tT)
for env in [{}, {"MXB_TUNE_CTAS_PER_SM": 4}, {"MXB_TUNE_CTAS_PER_SM": 8}, {"MXB_TUNE_CTAS_PER_SM": 16}, {"MXB_TUNE_CTAS_PER_SM": 32},
            {"MXB_TUNE_U": 1}, {"MXB_TUNE_U": 1, "MXB_TUNE_CTAS_PER_SM": 8}, {"MXB_TUNE_U": 4, "MXB_TUNE_CTAS_PER_SM": 8},
            {"MXB_TUNE_BLOCK": 128, "MXB_TUNE_CTAS_PER_SM": 16}, {"MXB_TUNE_BLOCK": 512, "MXB_TUNE_CTAS_PER_SM": 4}]:
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        ms, best = bc._time(ex, lambda: to.set(expr).run(ex), iters=6, warm=2)
        print(json.dumps({"env": env, "ms": round(ms, 4), "best": round(best, 4), "GBps": round(6 * n * 4 / ms / 1e6, 1), "kernel": ex.last_kernel()}), flush=True)
    except Exception as exc:
        print(json.dumps({"env": env, "error": str(exc)[:200]}), flush=True)
    for k in env:
        os.environ.pop(k)
