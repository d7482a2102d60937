"""Kernel-level timing of the other BASELINE.json configs (1, 3, 4, 5) at their full sizes; bench.py reports them
under `other_configs`.  CUDA events on the launching stream, working sets far larger than L2 (no flush needed)."""
from __future__ import annotations

from . import ops as mx


def _time(ex, fn, iters=8, warm=2):
    """(mean, best) device time of one launch in ms: three bursts of `iters` back-to-back launches, each between two
    events on the launching stream (an event pair around a single launch would also time the host-side lowering of
    the NEXT statement whenever the kernel is shorter than that, ~40 us)."""
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    bursts = []
    for _ in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        bursts.append(a.elapsed_time(b) / iters)
    return sum(bursts) / len(bursts), min(bursts)


def _entry(ex, ms, best, nbytes, nelem, peak):
    return {"ms_avg": ms, "ms_best": best, "GBps": nbytes / (ms * 1e-3) / 1e9, "frac_of_measured_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
            "Gelem_per_s": nelem / (ms * 1e-3) / 1e9, "algorithmic_bytes": nbytes, "kernel": ex.last_kernel()}


def run_c1(ex, peak):
    import torch
    rows, cols = 16384, 4096
    a, b = torch.rand(rows, cols, device="cuda"), torch.rand(rows, cols, device="cuda")
    c = torch.rand(rows, cols, device="cuda") - 0.5
    out = torch.empty(rows, device="cuda")
    ta, tb, tc, to = (mx.make_tensor(t) for t in (a, b, c, out))
    ms, best = _time(ex, lambda: to.set(mx.sum(ta * tb + tc, [1])).run(ex))
    ref = (a.double() * b.double() + c.double()).sum(1)
    err = ((out.double() - ref).abs() / ref.abs()).max().item()
    r = _entry(ex, ms, best, 3 * rows * cols * 4 + rows * 4, rows * cols, peak)
    r["max_rel_err_vs_fp64"] = err
    return {"sum(a*b+c,{1}) fp32 16384x4096": r}


def run_c3(ex, peak):
    import torch
    rows, cols = 65536, 8192
    x = torch.view_as_complex(torch.randn(rows, cols, 2, device="cuda"))
    tx = mx.make_tensor(x)
    n = rows * cols
    res = {}
    om = torch.empty(rows, dtype=torch.complex64, device="cuda")
    ms, best = _time(ex, lambda: mx.make_tensor(om).set(mx.mean(tx, [1])).run(ex), iters=5)
    res["mean(x,{1}) c64 65536x8192"] = _entry(ex, ms, best, n * 8 + rows * 8, n, peak)
    ov = torch.empty(rows, device="cuda")
    ms, best = _time(ex, lambda: mx.make_tensor(ov).set(mx.var(tx, [1], 1)).run(ex), iters=5)
    res["var(x,{1},ddof=1) c64 65536x8192"] = _entry(ex, ms, best, n * 8 + rows * 4, n, peak)
    oa, oi = torch.empty(rows, device="cuda"), torch.empty(rows, dtype=torch.int64, device="cuda")
    ms, best = _time(ex, lambda: mx.mtie(mx.make_tensor(oa), mx.make_tensor(oi)).set(mx.argmax(mx.abs2(tx), [1])).run(ex), iters=5)
    res["argmax(abs2(x),{1}) c64 65536x8192"] = _entry(ex, ms, best, n * 8 + rows * 12, n, peak)
    # spot checks on a few rows against torch (fp64)
    xs = x[:64].to(torch.complex128)
    res["mean(x,{1}) c64 65536x8192"]["max_abs_err_rows0_63"] = (om[:64].to(torch.complex128) - xs.mean(1)).abs().max().item()
    v64 = ((xs - xs.mean(1, keepdim=True)).abs() ** 2).sum(1) / (cols - 1)
    res["var(x,{1},ddof=1) c64 65536x8192"]["max_rel_err_rows0_63"] = ((ov[:64].double() - v64).abs() / v64).max().item()
    a2 = (x[:64].real.double() ** 2 + x[:64].imag.double() ** 2)
    res["argmax(abs2(x),{1}) c64 65536x8192"]["index_match_rows0_63"] = bool(
        ((oi[:64] - torch.arange(64, device="cuda") * cols) == a2.argmax(1)).all().item())
    return res


def black_scholes_expr(K, S, V, r, T):
    VsqrtT = V * mx.sqrt(T)
    d1 = (mx.log(S / K) + (r + 0.5 * V * V) * T) / VsqrtT
    d2 = d1 - VsqrtT
    return S * mx.normcdf(d1) - K * mx.exp(-1.0 * r * T) * mx.normcdf(d2)


def run_c4(ex, peak):
    import torch
    n = 1 << 28
    S = torch.rand(n, device="cuda") * 90 + 10
    K = torch.rand(n, device="cuda") * 90 + 10
    V = torch.rand(n, device="cuda") * 0.45 + 0.05
    r = torch.rand(n, device="cuda") * 0.09 + 0.01
    T = torch.rand(n, device="cuda") * 1.9 + 0.1
    out = torch.empty(n, device="cuda")
    tK, tS, tV, tr, tT, to = (mx.make_tensor(t) for t in (K, S, V, r, T, out))
    expr = black_scholes_expr(tK, tS, tV, tr, tT)
    ms, best = _time(ex, lambda: to.set(expr).run(ex), iters=5)
    res = _entry(ex, ms, best, 6 * n * 4, n, peak)
    m = 1 << 20
    s, k, v, rr, t = (z[:m].double() for z in (S, K, V, r, T))
    vs = v * t.sqrt()
    d1 = ((s / k).log() + (rr + 0.5 * v * v) * t) / vs
    d2 = d1 - vs
    N = lambda z: 0.5 * torch.erfc(-z / 2 ** 0.5)  # noqa: E731
    want = s * N(d1) - k * (-rr * t).exp() * N(d2)
    res["max_abs_err_first_2^20_vs_fp64"] = (out[:m].double() - want).abs().max().item()
    return {"black_scholes fp32 2^28 (5 in + 1 out)": res}


def run_c5(ex, peak):
    import torch
    d = 1024
    t = (torch.rand(d, d, d, device="cuda") * 0.25).to(torch.bfloat16)
    out = torch.empty(d, d, dtype=torch.bfloat16, device="cuda")
    tt, to = mx.make_tensor(t), mx.make_tensor(out)
    ms, best = _time(ex, lambda: to.set(mx.sum(mx.permute(tt, [2, 0, 1]), [2])).run(ex), iters=8)
    res = _entry(ex, ms, best, d * d * d * 2 + d * d * 2, d * d * d, peak)
    want = t[:8].float().sum(1).t()  # out[i][j] = sum_k t[j][k][i]
    res["max_rel_err_j0_7_vs_fp32"] = ((out[:, :8].float() - want).abs() / want).max().item()
    return {"sum(permute(t,{2,0,1}),{2}) bf16 1024^3": res}


def run_next(ex, peak):
    """SURVEY.md section 8f rows built after the five configs: the reference's own permute benchmark
    (bench/00_operators/operators.cu:40-59), a plain transpose, and cumsum (transforms/cub.h:2367-2395)."""
    import torch
    res = {}
    x = torch.randn(1000, 200, 6, 300, device="cuda")
    y = torch.empty(300, 1000, 6, 200, device="cuda")
    tx, ty = mx.make_tensor(x), mx.make_tensor(y)
    ms, best = _time(ex, lambda: ty.set(tx.Permute([3, 0, 2, 1])).run(ex))
    r = _entry(ex, ms, best, 2 * x.numel() * 4, x.numel(), peak)
    r["bit_exact"] = bool(torch.equal(y, x.permute(3, 0, 2, 1)))
    res["y = x.Permute({3,0,2,1}) fp32 {1000,200,6,300} (reference bench shape)"] = r
    del x, y
    a = torch.randn(8192, 8192, device="cuda")
    t = torch.empty(8192, 8192, device="cuda")
    ta, tt = mx.make_tensor(a), mx.make_tensor(t)
    ms, best = _time(ex, lambda: tt.set(ta.Permute([1, 0])).run(ex))
    r = _entry(ex, ms, best, 2 * a.numel() * 4, a.numel(), peak)
    r["bit_exact"] = bool(torch.equal(t, a.t()))
    res["t = a.Permute({1,0}) fp32 8192x8192"] = r
    del a, t
    torch.cuda.empty_cache()
    for name, shape in (("cumsum(x) fp32 16384x4096", (16384, 4096)), ("cumsum(x) fp32 2^28 (one row)", (1 << 28,))):
        x = torch.rand(*shape, device="cuda")
        y = torch.empty_like(x)
        tx, ty = mx.make_tensor(x), mx.make_tensor(y)
        ms, best = _time(ex, lambda: ty.set(mx.cumsum(tx)).run(ex))
        r = _entry(ex, ms, best, 2 * x.numel() * 4, x.numel(), peak)
        m = 1 << 16
        head = x.reshape(-1)[:m] if len(shape) == 1 else x[0, :m]
        got = y.reshape(-1)[:m] if len(shape) == 1 else y[0, :m]
        ref = head.double().cumsum(0)
        r["max_rel_err_first_row_head_vs_fp64"] = ((got.double() - ref).abs() / ref).max().item()
        res[name] = r
        del x, y
        torch.cuda.empty_cache()
    # variance where the row cannot stay on chip (one-pass op): a full tensor, 256 KB rows, column variance
    for name, shape, dims in (("var(x) fp32 2^30 (full tensor, one read)", (1 << 30,), None),
                              ("var(x,{1}) fp32 4096x65536 (256 KB rows)", (4096, 65536), [1]),
                              ("var(x,{0}) fp32 4096x65536 (column variance)", (4096, 65536), [0])):
        x = torch.rand(*shape, device="cuda") + 0.5
        oshape = () if dims is None else ((shape[0],) if dims == [1] else (shape[1],))
        o = torch.empty(oshape, device="cuda")
        tx, to = mx.make_tensor(x), mx.make_tensor(o)
        ms, best = _time(ex, lambda: to.set(mx.var(tx, dims, 1)).run(ex))
        r = _entry(ex, ms, best, x.numel() * 4 + o.numel() * 4, x.numel(), peak)
        if dims is None:
            want = x.double().var(unbiased=True)
            r["rel_err_vs_fp64"] = ((o.double() - want).abs() / want).item()
        else:
            xs = x[:64].double() if dims == [1] else x[:, :64].double()
            want = xs.var(dim=1 if dims == [1] else 0, unbiased=True)
            got = o[:64].double()
            r["max_rel_err_first_64_vs_fp64"] = ((got - want).abs() / want).max().item()
        res[name] = r
        del x, o
        torch.cuda.empty_cache()
    return res


BATCHED = ("C1 sum(a*b+c,{1}) fp32 16384x4096", "C3 mean(x,{1}) c64 65536x8192", "C3 var(x,{1},1) c64 65536x8192",
           "C3 argmax(abs2(x),{1}) c64 65536x8192", "C4 black_scholes fp32 2^28", "C5 sum(permute(t,{2,0,1}),{2}) bf16 1024^3")


def _time_burst(fn, use_graph, launches=10, reps=3):
    """Average device time of one launch: `launches` back-to-back calls captured into ONE CUDA graph (so the host-side
    lowering cannot starve an ~80 us kernel at 8 GPUs) and replayed `reps` times between two events on the current
    stream; falls back to plain back-to-back launches if the capture is refused."""
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    graph = None
    try:
        if not use_graph:   # the legacy default stream cannot be captured (N = 1: kernels of 0.1-1 ms need no graph)
            raise RuntimeError("no capture")
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
            for _ in range(launches):
                fn()
        graph = g
        graph.replay()
    except Exception:  # noqa: BLE001 - measured either way
        graph = None
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        if graph is not None:
            graph.replay()
        else:
            for _ in range(launches):
                fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / (reps * launches), graph is not None


def run_batched_shard(ex, rank, world, use_graph=False):
    """Configs 1, 3, 4, 5 sharded by their outermost batch dim over `world` ranks with NO communication:
    this rank times its own block.  Returns {name: {"ms", "bytes_total", "elems_total", "kernel", "graph"}} where the
    totals are those of the WHOLE job (all ranks), so total / max-over-ranks(ms) is the job's throughput."""
    import torch
    from .dist import shard_rows
    res = {}

    def _tb(fn):
        return _time_burst(fn, use_graph)

    def put(name, ms_graph, nbytes, nelem):
        res[name] = {"ms": ms_graph[0], "graph": ms_graph[1], "bytes_total": nbytes, "elems_total": nelem, "kernel": ex.last_kernel()}

    # C1
    rows, cols = 16384, 4096
    _, r = shard_rows(rows, rank, world)
    a, b = torch.rand(r, cols, device="cuda"), torch.rand(r, cols, device="cuda")
    c = torch.rand(r, cols, device="cuda") - 0.5
    out = torch.empty(r, device="cuda")
    ta, tb, tc, to = (mx.make_tensor(t) for t in (a, b, c, out))
    put(BATCHED[0], _tb(lambda: to.set(mx.sum(ta * tb + tc, [1])).run(ex)), 3 * rows * cols * 4 + rows * 4, rows * cols)
    del a, b, c, out
    # C3
    rows, cols = 65536, 8192
    _, r = shard_rows(rows, rank, world)
    x = torch.view_as_complex(torch.randn(r, cols, 2, device="cuda"))
    tx = mx.make_tensor(x)
    n = rows * cols
    om = torch.empty(r, dtype=torch.complex64, device="cuda")
    put(BATCHED[1], _tb(lambda: mx.make_tensor(om).set(mx.mean(tx, [1])).run(ex)), n * 8 + rows * 8, n)
    ov = torch.empty(r, device="cuda")
    put(BATCHED[2], _tb(lambda: mx.make_tensor(ov).set(mx.var(tx, [1], 1)).run(ex)), n * 8 + rows * 4, n)
    oa, oi = torch.empty(r, device="cuda"), torch.empty(r, dtype=torch.int64, device="cuda")
    put(BATCHED[3], _tb(lambda: mx.mtie(mx.make_tensor(oa), mx.make_tensor(oi)).set(mx.argmax(mx.abs2(tx), [1])).run(ex)),
        n * 8 + rows * 12, n)
    del x, tx, om, ov, oa, oi
    torch.cuda.empty_cache()
    # C4
    n = 1 << 28
    _, m = shard_rows(n, rank, world)
    S, K = torch.rand(m, device="cuda") * 90 + 10, torch.rand(m, device="cuda") * 90 + 10
    V, rr, T = torch.rand(m, device="cuda") * 0.45 + 0.05, torch.rand(m, device="cuda") * 0.09 + 0.01, torch.rand(m, device="cuda") * 1.9 + 0.1
    o4 = torch.empty(m, device="cuda")
    tK, tS, tV, tr, tT, to4 = (mx.make_tensor(t) for t in (K, S, V, rr, T, o4))
    expr = black_scholes_expr(tK, tS, tV, tr, tT)
    put(BATCHED[4], _tb(lambda: to4.set(expr).run(ex)), 6 * n * 4, n)
    del S, K, V, rr, T, o4
    torch.cuda.empty_cache()
    # C5: this rank owns whole 2 MiB slabs t[j] of the outermost dim and its columns out[:, j]
    d = 1024
    _, dj = shard_rows(d, rank, world)
    t = (torch.rand(dj, d, d, device="cuda") * 0.25).to(torch.bfloat16)
    o5 = torch.empty(d, dj, dtype=torch.bfloat16, device="cuda")
    tt, to5 = mx.make_tensor(t), mx.make_tensor(o5)
    put(BATCHED[5], _tb(lambda: to5.set(mx.sum(mx.permute(tt, [2, 0, 1]), [2])).run(ex)), d * d * d * 2 + d * d * 2, d * d * d)
    del t, o5
    torch.cuda.empty_cache()
    return res


def run_all(ex, peak):
    import torch
    out = {}
    for f in (run_c1, run_c3, run_c4, run_c5, run_next):
        try:
            out.update(f(ex, peak))
        except Exception as exc:  # report, do not hide
            out[f.__name__] = {"error": repr(exc)}
        torch.cuda.empty_cache()
    return out
