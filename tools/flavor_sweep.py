import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matx_b200 import bench_configs as bc, ops as mx
ex = mx.CudaExecutor()
n = 1 << 30
x = torch.rand(n, device="cuda"); tx = mx.make_tensor(x)
o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
for rep in range(2):
  for fl in [None, 0, 1, 2, 3, 4, 5]:
    if fl is None: os.environ.pop("MXB_LD_FLAVOR", None)
    else: os.environ["MXB_LD_FLAVOR"] = str(fl)
    try:
        ms, best = bc._time(ex, lambda: mx.make_tensor(o).set(mx.sum(tx)).run(ex), iters=10)
        ms2, best2 = bc._time(ex, lambda: mx.make_tensor(o).set(mx.max(tx)).run(ex), iters=10)
        print(json.dumps({"flavor": fl, "sum_ms": round(ms,4), "sum_best": round(best,4), "max_ms": round(ms2,4), "max_best": round(best2,4), "k": ex.last_kernel()[-12:]}), flush=True)
    except Exception as e:
        print(json.dumps({"flavor": fl, "error": str(e)[:300]}), flush=True)
# torch reference points
for name, f in (("torch.sum", lambda: x.sum()), ("torch.max", lambda: x.max())):
    ms, best = bc._time(ex, f, iters=10)
    print(json.dumps({"ref": name, "ms": round(ms,4), "best": round(best,4)}), flush=True)
