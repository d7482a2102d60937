"""Round-2 launch-shape sweep (development tool, run under gpurun): explicit (case, env) settings instead of a full grid.
    python tools/sweep2.py c2 c2small c3mean add      # prints one JSON line per setting
Every setting is a set of MXB_TUNE_* overrides read by the dispatcher; instances that are not ahead-of-time are JIT-built."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

PEAK = 6456.8


def run(name, nbytes, fn, envs, iters=10):
    for env in envs:
        env = {k: str(v) for k, v in env.items()}
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        try:
            ex = mx.CudaExecutor()   # a fresh handle per setting: the knobs are read when the handle is created
            ms, best = bc._time(ex, lambda: fn(ex), iters=iters, warm=3)
            print(json.dumps({"cfg": name, "env": env, "ms": round(ms, 4), "best": round(best, 4), "GBps": round(nbytes / ms / 1e6, 1),
                              "frac": round(nbytes / ms / 1e6 / PEAK, 3), "kernel": ex.last_kernel()}), flush=True)
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"cfg": name, "env": env, "error": str(exc)[:300]}), flush=True)
        finally:
            for k, v in old.items():
                if v is None:
                    os.environ.pop(k, None)
                else:
                    os.environ[k] = v


def dyn_envs():
    e = [{}, {"MXB_TUNE_DYNAMIC": 0}]
    for ct in (1, 2, 8, 16, 32):
        e.append({"MXB_TUNE_CHUNK_TILES": ct})
    e.append({"MXB_TUNE_BLOCK": 128})
    e.append({"MXB_TUNE_BLOCK": 128, "MXB_TUNE_CHUNK_TILES": 16})
    for v, u in ((8, 2), (8, 4)):
        for ct in (1, 2, 4, 16):
            e.append({"MXB_TUNE_V": v, "MXB_TUNE_U": u, "MXB_TUNE_CHUNK_TILES": ct})
    e.append({"MXB_TUNE_V": 8, "MXB_TUNE_U": 4, "MXB_TUNE_DYNAMIC": 0})
    e.append({"MXB_TUNE_V": 8, "MXB_TUNE_U": 4, "MXB_TUNE_BLOCK": 128, "MXB_TUNE_CHUNK_TILES": 4})
    e.append({"MXB_TUNE_V": 4, "MXB_TUNE_U": 8, "MXB_TUNE_CHUNK_TILES": 2})
    return e


which = sys.argv[1:] or ["c2", "c2small", "c3mean", "add"]
for tag, n in (("c2", 1 << 30), ("c2small", 1 << 27)):
    if tag not in which:
        continue
    x = torch.rand(n, device="cuda")
    tx = mx.make_tensor(x)
    o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    envs = dyn_envs()
    run(tag + ".sum", n * 4, lambda ex: mx.make_tensor(o).set(mx.sum(tx)).run(ex), envs)
    run(tag + ".argmax", n * 4, lambda ex: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx)).run(ex), envs[:9] + envs[13:18])
    del x, tx
    torch.cuda.empty_cache()
if "c3mean" in which:
    rows, cols = 65536, 8192
    x = torch.view_as_complex(torch.randn(rows, cols, 2, device="cuda"))
    tx = mx.make_tensor(x)
    om = torch.empty(rows, dtype=torch.complex64, device="cuda")
    envs = [{}, {"MXB_TUNE_DYNAMIC": 0}, {"MXB_TUNE_BLOCK": 128}, {"MXB_TUNE_BLOCK": 128, "MXB_TUNE_DYNAMIC": 0},
            {"MXB_TUNE_BLOCK": 64}, {"MXB_TUNE_BLOCK": 128, "MXB_TUNE_U": 4}, {"MXB_TUNE_BLOCK": 128, "MXB_TUNE_U": 1},
            {"MXB_TUNE_MINB": 6}, {"MXB_TUNE_MINB": 8}, {"MXB_TUNE_MINB": 8, "MXB_TUNE_BLOCK": 128, "MXB_TUNE_CTAS_PER_SM": 16},
            {"MXB_TUNE_V": 2, "MXB_TUNE_U": 4, "MXB_TUNE_BLOCK": 128}]
    run("c3.mean", rows * cols * 8, lambda ex: mx.make_tensor(om).set(mx.mean(tx, [1])).run(ex), envs)
    del x, tx
    torch.cuda.empty_cache()
if "add" in which:
    n = 1 << 28
    a, b, out = torch.rand(n, device="cuda"), torch.rand(n, device="cuda"), torch.empty(n, device="cuda")
    ta, tb, to = mx.make_tensor(a), mx.make_tensor(b), mx.make_tensor(out)
    envs = [{}]
    for v in (4, 8):
        for u in (1, 2, 4):
            envs.append({"MXB_TUNE_V": v, "MXB_TUNE_U": u})
    envs += [{"MXB_TUNE_V": 4, "MXB_TUNE_U": 1, "MXB_TUNE_BLOCK": 128}, {"MXB_TUNE_V": 4, "MXB_TUNE_U": 1, "MXB_TUNE_BLOCK": 512},
             {"MXB_TUNE_V": 4, "MXB_TUNE_U": 1, "MXB_LD_FLAVOR": 2}]
    run("vector_add", 3 * n * 4, lambda ex: to.set(ta + tb).run(ex), envs)
