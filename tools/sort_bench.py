"""sort timing (development tool, run under gpurun): one row of 2^24 / 2^26 fp32, 64 rows of 2^18, fp64, int32; torch.sort
beside it (CUB radix sort under torch) and a bit-exact comparison of the sorted values."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from matx_b200 import bench_configs as bc  # noqa: E402
from matx_b200 import ops as mx  # noqa: E402

ex = mx.CudaExecutor()
for shape, dt in (((1 << 24,), torch.float32), ((1 << 26,), torch.float32), ((64, 1 << 18), torch.float32), ((1 << 24,), torch.float64),
                  ((1 << 24,), torch.int32), ((100003,), torch.float32)):
    if dt.is_floating_point:
        x = torch.randn(shape, device="cuda", dtype=dt)
    else:
        x = torch.randint(-2 ** 31, 2 ** 31 - 1, shape, device="cuda", dtype=dt)
    out = torch.empty_like(x)
    for desc in (False, True):
        fn = lambda: mx.make_tensor(out).set(mx.sort(mx.make_tensor(x), mx.SORT_DIR_DESC if desc else mx.SORT_DIR_ASC)).run(ex)  # noqa: E731
        ms, best = bc._time(ex, fn, iters=5, warm=2)
        want = torch.sort(x, dim=-1, descending=desc).values
        ok = bool(torch.equal(out, want))
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            torch.sort(x, dim=-1, descending=desc)
        b.record()
        torch.cuda.synchronize()
        print(json.dumps({"shape": list(shape), "dtype": str(dt), "desc": desc, "ms": round(ms, 4), "torch_sort_ms": round(a.elapsed_time(b) / 5, 4), "equal": ok,
                          "Mkeys_per_s": round(x.numel() / ms / 1e3, 1), "kernel": ex.last_kernel()}), flush=True)
