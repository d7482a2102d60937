"""cumsum throughput (GB/s = bytes read + bytes written over time) across shapes and both grid modes, checked against
torch.cumsum on the device; torch's own cumsum timed beside it.  Run under gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from matx_b200 import ops as mx


def timed(f, iters=10):
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
    cases = [("f32 1 x 2^28", torch.float32, (1 << 28,)), ("f32 8 x 2^25", torch.float32, (8, 1 << 25)), ("f32 16384 x 4096", torch.float32, (16384, 4096)),
             ("f32 65536 x 1024", torch.float32, (65536, 1024)), ("f32 262144 x 256", torch.float32, (1 << 18, 256)), ("f32 1M x 64", torch.float32, (1 << 20, 64)), ("f32 300 x 100000", torch.float32, (300, 100000)),
             ("c64 65536 x 2048", torch.complex64, (65536, 2048)), ("f64 4096 x 16384", torch.float64, (4096, 16384)), ("bf16 16384 x 8192", torch.bfloat16, (16384, 8192)),
             ("i32 1 x 2^28", torch.int32, (1 << 28,))]
    for name, dt, shape in cases:
        if dt == torch.complex64:
            x = torch.view_as_complex(torch.rand(*shape, 2, device="cuda"))
        elif dt == torch.int32:
            x = torch.randint(-100, 100, shape, device="cuda", dtype=dt)
        else:
            x = torch.rand(*shape, device="cuda").to(dt)
        y = torch.empty_like(x)
        tx, ty = mx.make_tensor(x), mx.make_tensor(y)
        for mode in [0]:
            if mode:
                os.environ["MXB_SCAN_MODE"] = str(mode)
            ms = timed(lambda: ty.set(mx.cumsum(tx)).run(ex))
            os.environ.pop("MXB_SCAN_MODE", None)
            ref = torch.cumsum(x if dt != torch.bfloat16 else x.float(), dim=-1)
            if dt in (torch.int32,):
                ok = bool(torch.equal(y, ref.to(dt)))
                err = 0.0
            else:
                err = float(((y.to(ref.dtype) - ref).abs() / ref.abs().clamp_min(1e-6)).max())
                ok = err <= (1e-2 if dt == torch.bfloat16 else 1e-4)
            del ref
            ms_t = timed(lambda: torch.cumsum(x, dim=-1, out=y)) if dt != torch.complex64 else float("nan")
            nbytes = 2 * x.numel() * x.element_size()
            print(json.dumps({"case": name, "mode": {0: "auto", 1: "rows", 2: "tiles"}[mode], "ms": round(ms, 4), "GBps": round(nbytes / ms / 1e6, 1),
                              "frac_of_6456.8": round(nbytes / ms / 1e6 / 6456.8, 3), "torch_cumsum_GBps": round(nbytes / ms_t / 1e6, 1), "ok": ok,
                              "max_rel_vs_torch": err, "kernel": ex.last_kernel()}), flush=True)
        del x, y
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
