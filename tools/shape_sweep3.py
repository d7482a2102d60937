import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matx_b200 import bench_configs as bc, ops as mx
ex = mx.CudaExecutor()
def run(tag, fn, nbytes, envs):
    for env in envs:
        for k, v in env.items(): os.environ[k] = str(v)
        try:
            ms, _ = bc._time(ex, fn, iters=5)
            print(json.dumps({"case": tag, "env": env, "GBps": round(nbytes / ms / 1e6), "k": "/".join(ex.last_kernel().split("|")[0:1] + ex.last_kernel().split("|")[4:8])}), flush=True)
        except Exception as e:
            print(json.dumps({"case": tag, "env": env, "error": str(e)[:200]}), flush=True)
        for k in env: os.environ.pop(k)
envs = [{"MXB_TUNE_TEAM": 0}, {"MXB_TUNE_TEAM": 1}, {"MXB_TUNE_TEAM": 1, "MXB_TUNE_STEPS_PER_LANE": 8}]
for rows, cols in ((16384, 16384), (4096, 65536), (1024, 262144), (2000, 20000)):
    x = torch.rand(rows, cols, device="cuda"); tx = mx.make_tensor(x)
    o = torch.empty(rows, device="cuda"); oi = torch.empty(rows, dtype=torch.int64, device="cuda")
    run("sum %dx%d" % (rows, cols), lambda: mx.make_tensor(o).set(mx.sum(tx, [1])).run(ex), x.numel() * 4, envs)
    run("argmax %dx%d" % (rows, cols), lambda: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx, [1])).run(ex), x.numel() * 4, envs[:2])
    del x, tx; torch.cuda.empty_cache()
rows, cols = 65536, 8192
x = torch.view_as_complex(torch.randn(rows, cols, 2, device="cuda")); tx = mx.make_tensor(x)
om = torch.empty(rows, dtype=torch.complex64, device="cuda"); oa = torch.empty(rows, device="cuda"); oi = torch.empty(rows, dtype=torch.int64, device="cuda")
run("C3 mean", lambda: mx.make_tensor(om).set(mx.mean(tx, [1])).run(ex), x.numel() * 8, envs[:2])
run("C3 argmax(abs2)", lambda: mx.mtie(mx.make_tensor(oa), mx.make_tensor(oi)).set(mx.argmax(mx.abs2(tx), [1])).run(ex), x.numel() * 8, envs[:2])
