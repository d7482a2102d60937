"""Softmax shapes on one B200: GB/s of algorithmic bytes (one read + one write) per kernel family, torch.softmax beside it.
Usage: python tools/softmax_sweep.py  (prints one JSON line per shape)"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from matx_b200 import ops as mx  # noqa: E402


def timeit(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
    ev[0].record()
    for i in range(reps):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(reps))
    return ts[len(ts) // 2]


def main():
    ex = mx.CudaExecutor()
    cases = [((1, 10845, 8, 16), [3], torch.float32), ((1, 10845, 8, 16), None, torch.float32),
             ((4194304, 64), [1], torch.float32), ((1048576, 256), [1], torch.float32), ((262144, 1024), [1], torch.float32),
             ((65536, 4096), [1], torch.float32), ((16384, 16384), [1], torch.float32), ((4096, 65536), [1], torch.float32),
             ((64, 4194304), [1], torch.float32), ((1, 1 << 28), None, torch.float32), ((65536, 8192), [1], torch.bfloat16),
             ((1024, 512, 512), [1], torch.float32)]
    for shape, dims, dt in cases:
        x = torch.randn(shape, device="cuda", dtype=dt)
        o = torch.empty_like(x)
        st = mx.make_tensor(o).set(mx.softmax(mx.make_tensor(x), dims))
        n0 = ex.launch_count()
        st.run(ex)
        ex.sync()
        nl = ex.launch_count() - n0
        k = ex.last_kernel()
        if dims is None:
            ref = torch.softmax(x.flatten().float(), 0).reshape(shape)
        else:
            ref = torch.softmax(x.float(), dims[0])
        err = ((o.float() - ref).abs().max() / ref.abs().max()).item()
        t = timeit(lambda: st.run(ex))
        tdim = dims[0] if dims else None
        tt = timeit((lambda: torch.softmax(x, tdim)) if dims else (lambda: torch.softmax(x.flatten(), 0)))
        by = 2 * x.numel() * x.element_size()
        print(json.dumps({"shape": list(shape), "dims": dims, "dtype": str(dt).split(".")[-1], "ms": round(t, 4), "GBps": round(by / t / 1e6),
                          "launches": nl, "kernel": "/".join(k.split("|")[i] for i in (0, 4, 6)), "torch_ms": round(tt, 4),
                          "rel_err_vs_torch": float("%.2e" % err)}), flush=True)
        del x, o, ref
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
