"""find / find_idx over sizes, selectivities and dtypes: the single-pass look-back kernel (fast = 1) against the two-pass
count + scatter pair (MXB_SEL_TWO_PASS=1, fast = 0), with torch.masked_select beside it.  Development tool, run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import bench_configs as bc, ops as mx

ex = mx.CudaExecutor()
PEAK = 6456.8
for logn in (20, 24, 28):
    n = 1 << logn
    x = torch.rand(n, device="cuda")
    out = torch.empty(n, device="cuda")
    idx = torch.empty(n, dtype=torch.int32, device="cuda")
    nf = torch.zeros((), dtype=torch.int32, device="cuda")
    tx, to, ti, tn = (mx.make_tensor(t) for t in (x, out, idx, nf))
    for thr in (0.99, 0.5, 0.01):
        want = torch.masked_select(x, x > thr)
        ms_t, _ = bc._time(ex, lambda: torch.masked_select(x, x > thr), iters=5, warm=2)
        for fast in ("0", "1"):
            os.environ["MXB_SEL_TWO_PASS"] = "0" if fast == "1" else "1"
            for name, fn in (("find", lambda: mx.mtie(to, tn).set(mx.find(tx, mx.GT(thr))).run(ex)),
                             ("find_idx", lambda: mx.mtie(ti, tn).set(mx.find_idx(tx, mx.GT(thr))).run(ex))):
                try:
                    ms, best = bc._time(ex, fn, iters=6, warm=2)
                    ok = nf.item() == want.numel() and (torch.equal(out[: want.numel()], want) if name == "find"
                                                        else torch.equal(x[idx[: want.numel()].long()], want))
                    nbytes = n * 4 + want.numel() * 4
                    print(json.dumps({"n": n, "selected": round(want.numel() / n, 4), "op": name, "fast": int(fast), "ms": round(ms, 4),
                                      "GBps": round(nbytes / ms / 1e6, 1), "frac": round(nbytes / ms / 1e6 / PEAK, 3), "ok": bool(ok),
                                      "torch_masked_select_ms": round(ms_t, 4), "kernel": ex.last_kernel()}), flush=True)
                except Exception as exc:  # noqa: BLE001
                    print(json.dumps({"n": n, "op": name, "fast": int(fast), "error": str(exc)[:200]}), flush=True)
        os.environ.pop("MXB_SEL_TWO_PASS", None)
    del x, out, idx
    torch.cuda.empty_cache()
