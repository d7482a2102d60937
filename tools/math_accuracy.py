"""Accuracy of the hand-written fp32 device functions (mxb::f_normcdf, mxb::f_log) measured ON THE GPU against fp64
truth (scipy on the host), next to the CUDA library functions the reference calls (torch: special.ndtr / log, which
are normcdff-class library code).  Prints one JSON line per function; run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from scipy.special import ndtr

from matx_b200 import ops as mx


def ulps(got, ref64):
    ref32 = ref64.astype(np.float32)
    ulp = np.spacing(np.abs(ref32)).astype(np.float64)
    ok = np.isfinite(ref64) & (np.abs(ref64) > 1.2e-38)
    u = np.abs(got.astype(np.float64) - ref64)[ok] / ulp[ok]
    return u, ok


def report(name, x, got, ref64, extra=None):
    u, ok = ulps(got, ref64)
    i = int(np.argmax(u))
    rel = np.abs(got.astype(np.float64) - ref64)[ok] / np.abs(ref64[ok])
    d = {"fn": name, "n": int(x.size), "max_ulp": float(u.max()), "mean_ulp": float(u.mean()), "p999_ulp": float(np.quantile(u, 0.999)),
         "max_rel": float(rel.max()), "worst_x": float(x[ok][i])}
    if extra:
        d.update(extra)
    print(json.dumps(d), flush=True)


def main():
    ex = mx.CudaExecutor()
    rng = np.random.default_rng(0)
    n = 1 << 22
    # normcdf: the whole useful range, the centre, the deep tail
    for label, lo, hi in (("normcdf[-13,9]", -13.0, 9.0), ("normcdf[-1,1]", -1.0, 1.0), ("normcdf[-13,-5]", -13.0, -5.0)):
        x = rng.uniform(lo, hi, n).astype(np.float32)
        dx = torch.from_numpy(x).cuda()
        out = torch.empty_like(dx)
        mx.make_tensor(out).set(mx.normcdf(mx.make_tensor(dx))).run(ex)
        ex.sync()
        ref = ndtr(x.astype(np.float64))
        report(label, x, out.cpu().numpy(), ref, {"kernel": ex.last_kernel()})
        lib = torch.special.ndtr(dx).cpu().numpy()
        report(label + " (torch.special.ndtr)", x, lib, ref)
    spec = np.array([0.0, -0.0, np.inf, -np.inf, np.nan, 14.5, -14.5, 20.0, -20.0, -13.3, 1e-30, -1e-30], np.float32)
    ds = torch.from_numpy(spec).cuda()
    os_ = torch.empty_like(ds)
    mx.make_tensor(os_).set(mx.normcdf(mx.make_tensor(ds))).run(ex)
    ex.sync()
    print(json.dumps({"fn": "normcdf specials", "x": [repr(float(v)) for v in spec], "got": [repr(float(v)) for v in os_.cpu().numpy()]}), flush=True)

    for label, x in (("log exp(U[-80,80])", np.exp(rng.uniform(-80, 80, n)).astype(np.float32)),
                     ("log U[0.5,2]", rng.uniform(0.5, 2.0, n).astype(np.float32))):
        dx = torch.from_numpy(x).cuda()
        out = torch.empty_like(dx)
        mx.make_tensor(out).set(mx.log(mx.make_tensor(dx))).run(ex)
        ex.sync()
        ref = np.log(x.astype(np.float64))
        report(label, x, out.cpu().numpy(), ref)
        report(label + " (torch.log)", x, torch.log(dx).cpu().numpy(), ref)
    spec = np.array([0.0, -0.0, np.inf, -1.0, np.nan, 1e-40, 1.0, 3.4e38, 1.1754944e-38], np.float32)
    ds = torch.from_numpy(spec).cuda()
    os_ = torch.empty_like(ds)
    mx.make_tensor(os_).set(mx.log(mx.make_tensor(ds))).run(ex)
    ex.sync()
    want = np.log(spec.astype(np.float64)).astype(np.float32)
    print(json.dumps({"fn": "log specials", "x": [repr(float(v)) for v in spec], "got": [repr(float(v)) for v in os_.cpu().numpy()],
                      "want": [repr(float(v)) for v in want]}), flush=True)

    # Black-Scholes on the config-4 input distribution: ours vs fp64 truth vs the same chain in torch (library math)
    n = 1 << 22
    S, K = (rng.uniform(10, 100, n).astype(np.float32) for _ in range(2))
    V = rng.uniform(0.05, 0.5, n).astype(np.float32)
    r = rng.uniform(0.01, 0.1, n).astype(np.float32)
    T = rng.uniform(0.1, 2, n).astype(np.float32)
    d = {k: torch.from_numpy(v).cuda() for k, v in dict(S=S, K=K, V=V, r=r, T=T).items()}
    t = {k: mx.make_tensor(v) for k, v in d.items()}
    out = torch.empty(n, device="cuda")
    VsqrtT = t["V"] * mx.sqrt(t["T"])
    d1 = (mx.log(t["S"] / t["K"]) + (t["r"] + 0.5 * t["V"] * t["V"]) * t["T"]) / VsqrtT
    d2 = d1 - VsqrtT
    mx.make_tensor(out).set(t["S"] * mx.normcdf(d1) - t["K"] * mx.exp(-1.0 * t["r"] * t["T"]) * mx.normcdf(d2)).run(ex)
    ex.sync()
    S6, K6, V6, r6, T6 = (a.astype(np.float64) for a in (S, K, V, r, T))
    vs = V6 * np.sqrt(T6)
    e1 = (np.log(S6 / K6) + (r6 + 0.5 * V6 * V6) * T6) / vs
    truth = S6 * ndtr(e1) - K6 * np.exp(-r6 * T6) * ndtr(e1 - vs)
    tv = d["V"] * torch.sqrt(d["T"])
    td1 = (torch.log(d["S"] / d["K"]) + (d["r"] + 0.5 * d["V"] * d["V"]) * d["T"]) / tv
    lib = (d["S"] * torch.special.ndtr(td1) - d["K"] * torch.exp(-1.0 * d["r"] * d["T"]) * torch.special.ndtr(td1 - tv)).cpu().numpy()
    got = out.cpu().numpy()
    big = truth > 1.0
    print(json.dumps({"fn": "black_scholes", "n": n, "kernel": ex.last_kernel(),
                      "ours_max_abs": float(np.max(np.abs(got - truth))), "lib_max_abs": float(np.max(np.abs(lib - truth))),
                      "ours_max_rel_price>1": float(np.max(np.abs(got - truth)[big] / truth[big])),
                      "lib_max_rel_price>1": float(np.max(np.abs(lib - truth)[big] / truth[big])),
                      "ours_vs_lib_max_abs": float(np.max(np.abs(got - lib))),
                      "ours_max_abs_over_S": float(np.max(np.abs(got - truth) / S6)), "lib_max_abs_over_S": float(np.max(np.abs(lib - truth) / S6)),
                      "ours_vs_lib_max_abs_over_S": float(np.max(np.abs(got - lib) / S6))}), flush=True)


if __name__ == "__main__":
    main()
