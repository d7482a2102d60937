"""Permuted copy `(y = x.Permute(...)).run(exec)` — the reference's own benchmark shape (bench/00_operators/operators.cu:40-59:
x {1000,200,6,300} -> y {300,1000,6,200}, Permute({3,0,2,1})) plus plain 2-D transposes and a fused `a + permute(b)`;
GB/s = (bytes read + bytes written) / time.  Checks the result against torch on the device.  Run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import ops as mx


def timed(f, iters=20):
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        f()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


def main():
    ex = mx.CudaExecutor()
    cases = []
    for dt in (torch.float32, torch.float64, torch.complex64, torch.bfloat16):
        cases.append(("ref-bench {1000,200,6,300}.Permute({3,0,2,1})", dt, (1000, 200, 6, 300), (3, 0, 2, 1)))
    for dt in (torch.float32, torch.bfloat16, torch.complex64):
        cases.append(("transpose 8192x8192", dt, (8192, 8192), (1, 0)))
    cases.append(("transpose 10000x3001 (ragged)", torch.float32, (10000, 3001), (1, 0)))
    cases.append(("permute {2,0,1} 512x512x512", torch.float32, (512, 512, 512), (2, 0, 1)))
    cases.append(("permute {0,2,1} 64x1024x2048", torch.float32, (64, 1024, 2048), (0, 2, 1)))
    for name, dt, shape, perm in cases:
        if dt == torch.complex64:
            x = torch.view_as_complex(torch.randn(*shape, 2, device="cuda"))
        else:
            x = torch.randn(*shape, device="cuda").to(dt)
        y = torch.empty([shape[p] for p in perm], dtype=dt, device="cuda")
        tx, ty = mx.make_tensor(x), mx.make_tensor(y)
        f = lambda: ty.set(tx.Permute(list(perm))).run(ex)
        ms = timed(f)
        ok = bool(torch.equal(y, x.permute(*perm)))
        ms_torch = timed(lambda: y.copy_(x.permute(*perm)))
        nbytes = 2 * x.numel() * x.element_size()
        print(json.dumps({"case": name, "dtype": str(dt).replace("torch.", ""), "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
                          "torch_copy_GBps": round(nbytes / ms_torch / 1e6, 1), "exact": ok, "kernel": ex.last_kernel()}), flush=True)
    # fused: out = a + permute(b)  (one leaf walks rows, the other columns)
    a = torch.randn(8192, 8192, device="cuda")
    b = torch.randn(8192, 8192, device="cuda")
    o = torch.empty_like(a)
    ta, tb, to = mx.make_tensor(a), mx.make_tensor(b), mx.make_tensor(o)
    ms = timed(lambda: to.set(ta + tb.Permute([1, 0])).run(ex))
    ok = bool(torch.equal(o, a + b.t()))
    print(json.dumps({"case": "a + b.Permute({1,0}) 8192x8192", "dtype": "float32", "ms": round(ms, 4), "GBps": round(3 * a.numel() * 4 / ms / 1e6, 1),
                      "exact": ok, "kernel": ex.last_kernel()}), flush=True)


if __name__ == "__main__":
    main()
