"""One GPU's share of an 8-GPU run, on ONE GPU (development tool, run under gpurun): the batched configs through
bench_configs.run_batched_shard(rank 0 of 8) and the config-2 slab (2^27 fp32: sum, max, argmax), each under a list of
MXB_TUNE_* settings given as JSON objects on the command line."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

PEAK = 6456.8
envs = [{}] + [json.loads(a) for a in sys.argv[1:]]
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for env in envs:
        old = {k: os.environ.get(k) for k in env}
        os.environ.update(env)
        ex = mx.CudaExecutor(s)
        res = bc.run_batched_shard(ex, 0, 8, use_graph=True)
        for name, v in res.items():
            gb = v["bytes_total"] / 8 / v["ms"] / 1e6
            print(json.dumps({"cfg": name, "env": env, "ms": round(v["ms"], 5), "GBps": round(gb, 1), "frac": round(gb / PEAK, 3), "kernel": v["kernel"]}), flush=True)
        n = 1 << 27
        x = torch.rand(n, device="cuda")
        tx = mx.make_tensor(x)
        o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
        for name, fn in (("sum", lambda: mx.make_tensor(o).set(mx.sum(tx)).run(ex)), ("max", lambda: mx.make_tensor(o).set(mx.max(tx)).run(ex)),
                         ("argmax", lambda: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx)).run(ex))):
            ms, graph = bc._time_burst(fn, True)
            print(json.dumps({"cfg": "C2 slab 2^27 " + name, "env": env, "ms": round(ms, 5), "GBps": round(n * 4 / ms / 1e6, 1), "frac": round(n * 4 / ms / 1e6 / PEAK, 3),
                              "kernel": ex.last_kernel(), "graph": graph}), flush=True)
        del x, tx
        torch.cuda.empty_cache()
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
