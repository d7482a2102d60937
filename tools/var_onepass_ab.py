"""Variance A/B: the two-pass families (rows staged on chip: var_group / var_reg / var_tma / var_smem) against the
one-pass op (Welford + Chan through reduce_inner), time and accuracy against fp64 truth, over row lengths from 8
elements to a full tensor.  Development tool, run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import bench_configs as bc, ops as mx

ex = mx.CudaExecutor()
PEAK = 6456.8


def run(name, x, nbytes):
    rows = x.shape[0] if x.dim() == 2 else 1
    out = torch.empty(rows, device="cuda") if x.dim() == 2 else torch.empty((), device="cuda")
    tx, to = mx.make_tensor(x), mx.make_tensor(out)
    dims = [1] if x.dim() == 2 else None
    m = min(rows, 64)
    xs = (x[:m] if x.dim() == 2 else x[None, :]).to(torch.complex128 if x.is_complex() else torch.float64)
    truth = ((xs - xs.mean(1, keepdim=True)).abs() ** 2).sum(1) / (xs.shape[1] - 1)
    del xs
    for flag in ("0", "1"):
        os.environ["MXB_VAR_ONEPASS"] = flag
        os.environ["MXB_VAR_ONEPASS_MAX_R"] = "0"   # arm 0 = the two-pass families wherever the row fits on chip
        try:
            ms, best = bc._time(ex, lambda: to.set(mx.var(tx, dims, 1)).run(ex), iters=6, warm=2)
            got = (out[:m] if x.dim() == 2 else out[None]).double()
            err = ((got - truth).abs() / truth).max().item()
            print(json.dumps({"case": name, "onepass": int(flag), "ms": round(ms, 4), "best": round(best, 4), "GBps": round(nbytes / ms / 1e6, 1),
                              "frac": round(nbytes / ms / 1e6 / PEAK, 3), "max_rel_err_vs_fp64": err, "kernel": ex.last_kernel()}), flush=True)
        except Exception as exc:  # noqa: BLE001
            print(json.dumps({"case": name, "onepass": int(flag), "error": str(exc)[:200]}), flush=True)
    os.environ.pop("MXB_VAR_ONEPASS")
    os.environ.pop("MXB_VAR_ONEPASS_MAX_R")


x = torch.view_as_complex(torch.randn(65536, 8192, 2, device="cuda"))
run("C3 var c64 65536x8192", x, x.numel() * 8 + 65536 * 4)
del x
torch.cuda.empty_cache()
n = 1 << 28
for cols in (8, 32, 64, 256, 1000, 1024, 4096, 16384, 65536, 1 << 20):
    x = torch.rand(n // cols, cols, device="cuda") + 0.5
    run("f32 %dx%d" % (n // cols, cols), x, x.numel() * 4 + (n // cols) * 4)
    del x
    torch.cuda.empty_cache()
x = torch.rand(1 << 30, device="cuda") + 0.5
run("f32 full tensor 2^30", x, x.numel() * 4)
