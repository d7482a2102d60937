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
total = 1 << 28
for cols in (8, 64, 256, 4096, 8192):
    rows = total // cols
    x = torch.rand(rows, cols, device="cuda"); tx = mx.make_tensor(x)
    o = torch.empty(rows, device="cuda"); oi = torch.empty(rows, dtype=torch.int64, device="cuda")
    envs = [{}, {"MXB_TUNE_STEPS_PER_LANE": 1}, {"MXB_TUNE_STEPS_PER_LANE": 2}, {"MXB_TUNE_STEPS_PER_LANE": 8}] if cols <= 256 else [{}, {"MXB_TUNE_TEAM": 0}, {"MXB_TUNE_TEAM": 1}, {"MXB_TUNE_TEAM": 1, "MXB_TUNE_U": 2}]
    run("sum %dx%d" % (rows, cols), lambda: mx.make_tensor(o).set(mx.sum(tx, [1])).run(ex), total * 4, envs)
    run("argmax %dx%d" % (rows, cols), lambda: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx, [1])).run(ex), total * 4, envs[:3])
    if cols in (256, 2048):
        run("var %dx%d" % (rows, cols), lambda: mx.make_tensor(o).set(mx.var(tx, [1])).run(ex), total * 4, [{}, {"MXB_VAR_NO_GROUP": 1}, {"MXB_VAR_NO_GROUP": 1, "MXB_TUNE_VAR_VPT": 2}, {"MXB_VAR_NO_GROUP": 1, "MXB_TUNE_VAR_VPT": 8}])
    del x, tx; torch.cuda.empty_cache()
x = torch.rand(131072, 2048, device="cuda"); tx = mx.make_tensor(x); o = torch.empty(131072, device="cuda")
run("var 131072x2048", lambda: mx.make_tensor(o).set(mx.var(tx, [1])).run(ex), x.numel() * 4, [{}, {"MXB_TUNE_VAR_VPT": 2}, {"MXB_TUNE_VAR_VPT": 8}, {"MXB_TUNE_VAR_VPT": 1}])
del x, tx; torch.cuda.empty_cache()
rows, cols = 16384, 4096
a, b, c = (torch.rand(rows, cols, device="cuda") for _ in range(3)); out = torch.empty(rows, device="cuda")
ta, tb, tc, to = (mx.make_tensor(t) for t in (a, b, c, out))
run("C1", lambda: to.set(mx.sum(ta * tb + tc, [1])).run(ex), 3 * rows * cols * 4, [{}, {"MXB_TUNE_TEAM": 0}, {"MXB_TUNE_TEAM": 1}, {"MXB_TUNE_TEAM": 1, "MXB_TUNE_U": 4}])
