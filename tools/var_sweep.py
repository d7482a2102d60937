import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matx_b200 import bench_configs as bc, ops as mx
ex = mx.CudaExecutor()
rows, cols = 65536, 8192
x = torch.view_as_complex(torch.randn(rows, cols, 2, device="cuda"))
tx = mx.make_tensor(x); ov = torch.empty(rows, device="cuda")
n = rows*cols
for env in [{}, {"MXB_VAR_NO_TMA": 1}]:
    for k, v in env.items(): os.environ[k] = str(v)
    try:
        ms, best = bc._time(ex, lambda: mx.make_tensor(ov).set(mx.var(tx, [1], 1)).run(ex), iters=8)
        print(json.dumps({"env": env, "ms": round(ms, 4), "best": round(best, 4), "GBps": round(n*8/ms/1e6, 1), "kernel": ex.last_kernel()}), flush=True)
    except Exception as e:
        print(json.dumps({"env": env, "error": str(e)[:300]}), flush=True)
    for k in env: os.environ.pop(k)
xs = x[:64].to(torch.complex128)
v64 = ((xs - xs.mean(1, keepdim=True)).abs() ** 2).sum(1) / (cols - 1)
print("max rel err rows0-63:", ((ov[:64].double() - v64).abs() / v64).max().item())
# f32 rows of 32 KB and 100 KB
for cols2 in (2048, 4096, 8192, 16384, 25600):
    y = torch.rand(20000, cols2, device="cuda") + 1
    oy = torch.empty(20000, device="cuda")
    for envv in ({"MXB_VAR_NO_TMA": "1"}, {"MXB_TUNE_BLOCK": "256"}, {"MXB_TUNE_BLOCK": "512"}, {}):
        os.environ.update(envv)
        ms, best = bc._time(ex, lambda: mx.make_tensor(oy).set(mx.var(mx.make_tensor(y), [1], 1)).run(ex), iters=5)
        print(json.dumps({"f32 cols": cols2, "env": envv, "ms": round(ms, 4), "GBps": round(y.numel()*4/ms/1e6, 1), "kernel": ex.last_kernel()}), flush=True)
        for kk in envv: os.environ.pop(kk)
    ref = y[:32].double().var(1, unbiased=True)
    print(json.dumps({"f32 cols": cols2, "ms": round(ms, 4), "GBps": round(y.numel()*4/ms/1e6, 1), "kernel": ex.last_kernel(), "err": ((oy[:32].double()-ref).abs()/ref).max().item()}), flush=True)
