"""Launch-shape sweep over the headline kernels (development tool, run under gpurun).  Every setting is a set of
MXB_TUNE_* overrides; kernels that are not ahead-of-time are JIT-compiled on the fly.  Prints one line per setting."""
import itertools
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

ex = mx.CudaExecutor()
PEAK = 6456.8


def timed(fn, iters=6):
    ms, best = bc._time(ex, fn, iters=iters, warm=2)
    return ms, best


def with_env(env, fn):
    old = {k: os.environ.get(k) for k in env}
    for k, v in env.items():
        os.environ[k] = str(v)
    try:
        return fn()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def sweep(name, nbytes, fn, grid):
    keys = list(grid)
    for vals in itertools.product(*[grid[k] for k in keys]):
        env = {k: v for k, v in zip(keys, vals) if v is not None}
        try:
            ms, best = with_env(env, lambda: timed(fn))
            print(json.dumps({"cfg": name, "env": env, "ms": round(ms, 4), "best": round(best, 4), "GBps": round(nbytes / ms / 1e6, 1),
                              "frac": round(nbytes / ms / 1e6 / PEAK, 3), "kernel": ex.last_kernel()}), flush=True)
        except Exception as exc:
            print(json.dumps({"cfg": name, "env": env, "error": str(exc)[:200]}), flush=True)


which = sys.argv[1:] or ["c2", "c1", "c3", "c4", "c5"]
if "c2" in which:
    n = 1 << 30
    x = torch.rand(n, device="cuda")
    tx = mx.make_tensor(x)
    o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    g = {"MXB_TUNE_V": [8], "MXB_TUNE_U": [1, 2, 4], "MXB_TUNE_BLOCK": [256, 512], "MXB_TUNE_CTAS_PER_SM": [4, 8, 16]}
    sweep("c2.sum", n * 4, lambda: mx.make_tensor(o).set(mx.sum(tx)).run(ex), g)
    sweep("c2.argmax", n * 4, lambda: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx)).run(ex), g)
    del x, tx
    torch.cuda.empty_cache()
if "c1" in which:
    rows, cols = 16384, 4096
    a, b, c = (torch.rand(rows, cols, device="cuda") for _ in range(3))
    out = torch.empty(rows, device="cuda")
    ta, tb, tc, to = (mx.make_tensor(t) for t in (a, b, c, out))
    g = {"MXB_TUNE_V": [8], "MXB_TUNE_U": [1, 2], "MXB_TUNE_BLOCK": [128, 256], "MXB_TUNE_CTAS_PER_SM": [8, 16], "MXB_TUNE_TEAM": [0]}
    sweep("c1.fma_sum", 3 * rows * cols * 4, lambda: to.set(mx.sum(ta * tb + tc, [1])).run(ex), g)
    del a, b, c
    torch.cuda.empty_cache()
if "c3" in which:
    rows, cols = 65536, 8192
    x = torch.view_as_complex(torch.randn(rows, cols, 2, device="cuda"))
    tx = mx.make_tensor(x)
    n = rows * cols
    oa, oi = torch.empty(rows, device="cuda"), torch.empty(rows, dtype=torch.int64, device="cuda")
    ov = torch.empty(rows, device="cuda")
    g = {"MXB_TUNE_V": [4], "MXB_TUNE_U": [1, 2, 4], "MXB_TUNE_BLOCK": [256], "MXB_TUNE_CTAS_PER_SM": [8, 16]}
    om = torch.empty(rows, dtype=torch.complex64, device="cuda")
    sweep("c3.mean", n * 8, lambda: mx.make_tensor(om).set(mx.mean(tx, [1])).run(ex), g)
    sweep("c3.argmax_abs2", n * 8, lambda: mx.mtie(mx.make_tensor(oa), mx.make_tensor(oi)).set(mx.argmax(mx.abs2(tx), [1])).run(ex), g)
    g = {"MXB_VAR_SMEM_ONLY": [1], "MXB_TUNE_V": [4], "MXB_TUNE_U": [2, 4, 8], "MXB_TUNE_BLOCK": [512, 1024], "MXB_TUNE_CTAS_PER_SM": [16]}
    sweep("c3.var", n * 8, lambda: mx.make_tensor(ov).set(mx.var(tx, [1], 1)).run(ex), g)
    del x, tx
    torch.cuda.empty_cache()
if "c4" in which:
    n = 1 << 28
    S = torch.rand(n, device="cuda") * 90 + 10
    K = torch.rand(n, device="cuda") * 90 + 10
    V = torch.rand(n, device="cuda") * 0.45 + 0.05
    r = torch.rand(n, device="cuda") * 0.09 + 0.01
    T = torch.rand(n, device="cuda") * 1.9 + 0.1
    out = torch.empty(n, device="cuda")
    tK, tS, tV, tr, tT, to = (mx.make_tensor(t) for t in (K, S, V, r, T, out))
    expr = bc.black_scholes_expr(tK, tS, tV, tr, tT)
    g = {"MXB_TUNE_V": [8], "MXB_TUNE_U": [1], "MXB_TUNE_BLOCK": [128, 256], "MXB_TUNE_CTAS_PER_SM": [16, 32]}
    sweep("c4.black_scholes", 6 * n * 4, lambda: to.set(expr).run(ex), g)
    g = {"MXB_TUNE_V": [8], "MXB_TUNE_U": [1, 2, 4], "MXB_TUNE_BLOCK": [128, 256], "MXB_TUNE_CTAS_PER_SM": [32, 100000]}
    sweep("vector_add", 3 * n * 4, lambda: to.set(tS + tK).run(ex), g)
    del S, K, V, r, T, out
    torch.cuda.empty_cache()
if "c5" in which:
    d = 1024
    t = (torch.rand(d, d, d, device="cuda") * 0.25).to(torch.bfloat16)
    out = torch.empty(d, d, dtype=torch.bfloat16, device="cuda")
    tt, to = mx.make_tensor(t), mx.make_tensor(out)
    g = {"MXB_TUNE_V": [16], "MXB_TUNE_U": [2, 4], "MXB_TUNE_BLOCK": [256], "MXB_TUNE_TX": [16, 32, 64], "MXB_TUNE_CTAS_PER_SM": [8]}
    sweep("c5.bf16_permuted_sum", d * d * d * 2, lambda: to.set(mx.sum(mx.permute(tt, [2, 0, 1]), [2])).run(ex), g)
