import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from matx_b200 import bench_configs as bc, ops as mx
ex = mx.CudaExecutor()
big = torch.rand(1 << 30, device="cuda")   # keep L2 cold between iterations by walking different slabs
o, oi = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
for logn in (27, 26):
    n = 1 << logn
    slabs = [mx.make_tensor(big[i * n:(i + 1) * n]) for i in range((1 << 30) // n)]
    for env in [{}, {"MXB_TUNE_CTAS_PER_SM": 4}, {"MXB_TUNE_CTAS_PER_SM": 6}, {"MXB_TUNE_CTAS_PER_SM": 12}, {"MXB_TUNE_CTAS_PER_SM": 16},
                {"MXB_TUNE_BLOCK": 512, "MXB_TUNE_CTAS_PER_SM": 4}, {"MXB_TUNE_BLOCK": 512, "MXB_TUNE_CTAS_PER_SM": 2}, {"MXB_TUNE_U": 2}, {"MXB_TUNE_U": 8, "MXB_TUNE_CTAS_PER_SM": 4}]:
        for k, v in env.items(): os.environ[k] = str(v)
        res = {}
        for name in ("sum", "argmax"):
            it = [0]
            def f():
                t = slabs[it[0] % len(slabs)]; it[0] += 1
                if name == "sum": mx.make_tensor(o).set(mx.sum(t)).run(ex)
                else: mx.mtie(mx.make_tensor(o), mx.make_tensor(oi)).set(mx.argmax(t)).run(ex)
            # back-to-back launches: measures the steady per-kernel cost including launch gaps
            f(); f(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(64): f()
            b.record(); torch.cuda.synchronize()
            res[name] = round(a.elapsed_time(b) / 64 * 1000, 2)
        print(json.dumps({"log2n": logn, "env": env, "us_per_kernel": res, "ideal_us@7.2TB/s": round(n * 4 / 7.2e6, 1)}), flush=True)
        for k in env: os.environ.pop(k)
