"""Bandwidth of batched reductions over a range of row lengths / layouts (development tool; finds weak launch shapes)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matx_b200 import bench_configs as bc, ops as mx
ex = mx.CudaExecutor()
total = 1 << 28   # elements (1 GiB fp32)
for cols in (8, 32, 64, 128, 256, 512, 1000, 1024, 2048, 4096, 16384, 65536, 1 << 20):
    rows = total // cols
    x = torch.rand(rows, cols, device="cuda")
    tx = mx.make_tensor(x)
    res = {"shape": [rows, cols]}
    o = torch.empty(rows, device="cuda"); oi = torch.empty(rows, dtype=torch.int64, device="cuda")
    for name, fn in (("sum", lambda: mx.make_tensor(o).set(mx.sum(tx, [1])).run(ex)),
                     ("argmax", lambda: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(tx, [1])).run(ex)),
                     ("var", lambda: mx.make_tensor(o).set(mx.var(tx, [1])).run(ex))):
        ms, _ = bc._time(ex, fn, iters=5)
        res[name] = [round(rows * cols * 4 / ms / 1e6), ex.last_kernel().split("|")[0] + "/" + "/".join(ex.last_kernel().split("|")[4:7])]
    oc = torch.empty(cols, device="cuda")
    ms, _ = bc._time(ex, lambda: mx.make_tensor(oc).set(mx.sum(tx, [0])).run(ex), iters=5)
    res["sum_dim0"] = [round(rows * cols * 4 / ms / 1e6), ex.last_kernel().split("|")[0]]
    ms, _ = bc._time(ex, lambda: x.sum(1), iters=5)
    res["torch_sum"] = round(rows * cols * 4 / ms / 1e6)
    print(json.dumps(res), flush=True)
    del x, tx
    torch.cuda.empty_cache()
