"""-m gpu parity tests of cumsum (mxb_cumsum, the `scan` kernel family) against the CPU oracle's sequential running sum
(std::partial_sum, what the reference's HostExecutor does) and numpy.  Reference tests mirrored: test/00_tensor/
CUBTests.cu:203-226 (permuted integer input), :536-575 (batched float rows), :640-670 (complex rows)."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import np_tensor, f32_to_bf16_bits, bf16_bits_to_f32

pytestmark = pytest.mark.gpu


def run_cumsum(oracle, build, arrays, out_shape, out_dtype, dtypes=None, env=None):
    import os
    import torch
    dtypes = dtypes or [None] * len(arrays)
    dev = [G.to_dev(a, d) for a, d in zip(arrays, dtypes)]
    npdt = G._NP_OF[out_dtype]
    tdt = {A.BF16: torch.bfloat16, A.F16: torch.float16}.get(out_dtype)
    out_d = torch.zeros(out_shape, dtype=tdt, device="cuda") if tdt else torch.from_numpy(np.zeros(out_shape, npdt)).cuda()
    ex = G.executor()
    old = {k: os.environ.get(k) for k in (env or {})}
    os.environ.update({k: str(v) for k, v in (env or {}).items()})
    try:
        mx.make_tensor(out_d).set(mx.cumsum(build(*[mx.make_tensor(t) for t in dev]))).run(ex)
        ex.sync()
    finally:
        for k, v in old.items():
            os.environ.pop(k, None) if v is None else os.environ.__setitem__(k, v)
    k = ex.last_kernel()
    got = G.from_dev(out_d, out_dtype)
    want = np.zeros(out_shape, npdt)
    oracle.cumsum(mx.cumsum(build(*[np_tensor(np.ascontiguousarray(a), d) for a, d in zip(arrays, dtypes)])), want, out_dtype)
    return got, want, k


def test_reference_known_answers(oracle):
    inv = np.array([[1, 2, 3, 4], [10, 20, 30, 40], [100, 200, 300, 400]], np.int32)
    got, want, k = run_cumsum(oracle, lambda t: t.Permute([1, 0]), [inv], (4, 3), A.I32)
    assert k.startswith("scan|"), k
    assert np.array_equal(got, np.cumsum(inv.T, axis=1)) and np.array_equal(got, want)
    a = np.array([1, 2, 3, 4, 5], np.float32)
    got, want, _ = run_cumsum(oracle, lambda t: t, [a], (5,), A.F32)
    assert got.tolist() == [1, 3, 6, 10, 15]
    x = np.zeros((2, 3, 40), np.float32)
    for i in range(2):
        for j in range(3):
            x[i, j] = 1000 * i + 100 * j + (40 - np.arange(40))
    got, want, _ = run_cumsum(oracle, lambda t: t, [x], x.shape, A.F32)
    assert np.max(np.abs(got - np.cumsum(x.astype(np.float64), axis=2))) <= 0.001   # the reference's own bar


@pytest.mark.parametrize("shape", [(1,), (7,), (1023,), (1024,), (1025,), (4096,), (4097,), (20000,), (3, 5000), (300, 257), (2000, 64),
                                   (5, 3, 1000), (1, 1 << 20), (2, 300001), (5000, 100), (1000, 2048), (70000, 7), (600, 513),
                                   (100000, 8), (50000, 16), (3000, 33), (40000, 4), (9999, 1), (7777, 128), (1201, 124)])
def test_int32_exact_all_shapes(oracle, shape):
    """Integer sums do not care about the order of additions: every tiling / carry path must be bit-exact."""
    rng = np.random.default_rng(sum(shape))
    x = rng.integers(-1000, 1000, shape).astype(np.int32)
    got, want, k = run_cumsum(oracle, lambda t: t, [x], shape, A.I32)
    assert k.startswith("scan|"), k
    assert np.array_equal(got, want)
    assert np.array_equal(got, np.cumsum(x.astype(np.int64), axis=-1).astype(np.int32))


@pytest.mark.parametrize("mode", [1, 2])     # 1 = a CTA walks whole rows with a register carry, 2 = one CTA per tile + exchange
def test_both_grid_modes_exact(oracle, mode):
    rng = np.random.default_rng(50 + mode)
    for shape in [(4, 70000), (1, 1 << 21), (40, 9000)]:
        x = rng.integers(-50, 50, shape).astype(np.int32)
        got, want, k = run_cumsum(oracle, lambda t: t, [x], shape, A.I32, env={"MXB_SCAN_MODE": mode})
        assert np.array_equal(got, want), (mode, shape)


def test_warp_and_cta_teams_agree_exactly(oracle):
    rng = np.random.default_rng(55)
    for shape in [(700, 64), (900, 1000), (3, 129)]:
        x = rng.integers(-50, 50, shape).astype(np.int32)
        for team in (0, 1):
            got, want, k = run_cumsum(oracle, lambda t: t, [x], shape, A.I32, env={"MXB_TUNE_TEAM": team})
            assert ("|T%d|" % team) in k + "|", k
            assert np.array_equal(got, want), (team, shape)
    xf = rng.random((800, 300)).astype(np.float32)
    for team in (0, 1):
        got, want, k = run_cumsum(oracle, lambda t: t, [xf], xf.shape, A.F32, env={"MXB_TUNE_TEAM": team})
        truth = np.cumsum(xf.astype(np.float64), axis=1)
        assert np.max(np.abs(got - truth) / truth) <= 1e-5


@pytest.mark.parametrize("dt,tol", [(A.F32, 1e-5), (A.F64, 1e-12), (A.C64, 1e-5)])
def test_floating_within_tolerance(oracle, dt, tol):
    rng = np.random.default_rng(60)
    for shape in [(33,), (10000,), (64, 3000), (2, 1 << 19)]:
        if dt == A.C64:
            x = (rng.random(shape) + 1j * rng.random(shape)).astype(np.complex64)
        else:
            x = rng.random(shape).astype(np.float32 if dt == A.F32 else np.float64)   # positive: no cancellation
        got, want, _ = run_cumsum(oracle, lambda t: t, [x], shape, dt)
        truth = np.cumsum(x.astype(np.complex128 if dt == A.C64 else np.float64), axis=-1)
        # tolerance stated on each prefix: |got - truth| <= tol * |truth| (north star: 1e-5 relative for fp32)
        assert np.max(np.abs(got - truth) / np.abs(truth)) <= tol
        # the sequential fp32 running sum of the reference's host path is itself ~sqrt(n) ulp off; both agree with truth
        assert np.max(np.abs(want - truth) / np.abs(truth)) <= max(tol, 2e-4)


def test_run_to_run_deterministic():
    import torch
    ex = G.executor()
    x = torch.randn(3, 1 << 20, device="cuda")
    a, b = torch.empty_like(x), torch.empty_like(x)
    mx.make_tensor(a).set(mx.cumsum(mx.make_tensor(x))).run(ex)
    for _ in range(5):
        mx.make_tensor(b).set(mx.cumsum(mx.make_tensor(x))).run(ex)
        ex.sync()
        assert torch.equal(a, b)


def test_fused_strided_and_bf16(oracle):
    rng = np.random.default_rng(61)
    a = rng.random((50, 700)).astype(np.float32)
    b = rng.random((50, 700)).astype(np.float32)
    got, want, k = run_cumsum(oracle, lambda x, y: x * y + 1.0, [a, b], a.shape, A.F32)
    assert np.max(np.abs(got - np.cumsum(a.astype(np.float64) * b + 1, axis=1)) / np.cumsum(a.astype(np.float64) * b + 1, axis=1)) <= 1e-5
    # scan along a strided dim (permuted view): the scalar instance
    t = rng.integers(-9, 9, (300, 40)).astype(np.int32)
    got, want, k = run_cumsum(oracle, lambda x: x.Permute([1, 0]), [t], (40, 300), A.I32)
    assert "|V1|" in k, k
    assert np.array_equal(got, np.cumsum(t.T, axis=1))
    # bf16 in, fp32 accumulation, one rounding at the store
    h = f32_to_bf16_bits(rng.random((8, 3000)).astype(np.float32))
    got, want, k = run_cumsum(oracle, lambda x: x, [h], h.shape, A.BF16, dtypes=[A.BF16])
    truth = np.cumsum(bf16_bits_to_f32(h).astype(np.float64), axis=1)
    assert np.max(np.abs(bf16_bits_to_f32(got) - truth) / truth) <= 2 ** -8


def test_full_size_properties():
    """2^28 fp32 elements in one row (the tile-exchange mode at scale): last element = the sum; differences give the input back."""
    import torch
    ex = G.executor()
    n = 1 << 28
    x = (torch.rand(n, device="cuda") * 2 - 1).round() * 3      # integers in {-3, 0, 3}: every prefix is exact in fp32
    y = torch.empty_like(x)
    mx.make_tensor(y).set(mx.cumsum(mx.make_tensor(x))).run(ex)
    ex.sync()
    assert ex.last_kernel().startswith("scan|")
    assert float(y[-1]) == float(x.double().sum())
    assert torch.equal(y[1:] - y[:-1], x[1:]) and float(y[0]) == float(x[0])


def test_two_streams_run_long_row_cumsum_concurrently():
    """the TILES-mode scan spins on tiles of other CTAs, so every CTA of its grid must be resident.  It is launched
    cooperatively (the runtime guarantees co-residency or refuses), so two handles scanning long rows on two streams at
    once must both finish — and be right."""
    import threading

    import torch
    n = 1 << 24
    xs = [torch.rand(n, device="cuda") for _ in range(2)]
    outs = [torch.empty(n, device="cuda") for _ in range(2)]
    streams = [torch.cuda.Stream() for _ in range(2)]
    exs = [mx.CudaExecutor(s) for s in streams]
    torch.cuda.synchronize()

    def work(i):
        with torch.cuda.stream(streams[i]):
            for _ in range(20):
                mx.make_tensor(outs[i]).set(mx.cumsum(mx.make_tensor(xs[i]))).run(exs[i])
        exs[i].sync()
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts), "a concurrent long-row cumsum did not finish"
    for x, o in zip(xs, outs):
        want = torch.cumsum(x.double(), 0)
        assert float(((o.double() - want).abs() / want).max().item()) <= 1e-5


def test_one_executor_shared_by_two_host_threads():
    """SURVEY 8b threading: one handle used from two host threads (a b200Executor copied into two threads shares it): the
    entry points lock the handle, the statements serialise on its stream, every result is right."""
    import threading

    import torch
    ex = mx.CudaExecutor()
    n = 1 << 22
    xs = [torch.rand(n, device="cuda") + i for i in range(2)]
    res = [[], []]

    def work(i):
        o = torch.zeros((), device="cuda")
        m = torch.zeros((), device="cuda")
        idx = torch.zeros((), dtype=torch.int64, device="cuda")
        for _ in range(50):
            mx.make_tensor(o).set(mx.sum(mx.make_tensor(xs[i]))).run(ex)
            mx.mtie(mx.make_tensor(m), mx.make_tensor(idx)).set(mx.argmax(mx.make_tensor(xs[i]))).run(ex)
        ex.sync()
        res[i] = [o.item(), m.item(), idx.item()]
    ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
    for t in ts:
        t.start()
    for t in ts:
        t.join(timeout=120)
    assert not any(t.is_alive() for t in ts)
    for i in range(2):
        truth = xs[i].double().sum().item()
        assert abs(res[i][0] - truth) <= 1e-5 * truth
        assert res[i][1] == xs[i].max().item() and res[i][2] == int(torch.argmax(xs[i]).item())


# ---- warp tiles (round 2, the default for few long rows): tile = 1024 elements (512 for 8-byte values), a group is 32
# tiles, a supergroup 1024 tiles; sizes on and around every edge, several rows with ragged ends, all exchange paths ----
@pytest.mark.parametrize("shape", [(32767,), (32768,), (32769,), ((1 << 20) - 1,), (1 << 20,), ((1 << 20) + 1,), (3 * (1 << 20) + 1025,),
                                   (3, (1 << 20) + 77), (5, 40001), (2, 2 * (1 << 20))])
def test_warp_tile_scan_edges_int32_exact(oracle, shape):
    rng = np.random.default_rng(sum(shape))
    x = rng.integers(-1000, 1000, shape).astype(np.int32)
    for env in ({}, {"MXB_SCAN_WTILES": 0}):         # warp tiles, and the flat CTA-tile exchange kept beside them
        got, want, k = run_cumsum(oracle, lambda t: t, [x], shape, A.I32, env=env)
        assert k.startswith("scan|"), k
        assert np.array_equal(got, np.cumsum(x.astype(np.int64), axis=-1).astype(np.int32)), (shape, env)


@pytest.mark.parametrize("dt,npdt,tol", [(A.F32, np.float32, 1e-5), (A.F64, np.float64, 1e-12), (A.I64, np.int64, 0)])
def test_warp_tile_scan_value_widths_and_operand_kinds(oracle, dt, npdt, tol):
    """8-byte values take 512-element tiles; an unaligned row start takes the scalar loads; an expression operand the
    generic loads (no L2 hints); a plain contiguous tensor the hinted 16-byte loads."""
    rng = np.random.default_rng(dt)
    n = (1 << 20) + 513
    x = (rng.random(n + 1) * 4).astype(npdt) if tol else rng.integers(-9, 9, n + 1).astype(npdt)
    for view_of, arr, ref in ((lambda t: t, x[:n].copy(), x[:n]), (lambda t: t.Slice([1], [n + 1]), x, x[1:]), (lambda t: t + t, x[:n].copy(), x[:n] + x[:n])):
        got, want, k = run_cumsum(oracle, view_of, [arr], (n,), dt)
        truth = np.cumsum(ref.astype(np.float64 if tol else np.int64))
        if tol:
            assert np.max(np.abs(got.astype(np.float64) - truth) / np.maximum(np.abs(truth), 1.0)) <= tol, (dt, k)
        else:
            assert np.array_equal(got, truth.astype(npdt)), (dt, k)


def test_warp_tile_scan_is_deterministic_and_graph_replayable():
    import torch
    x = torch.rand((1 << 22) + 5, device="cuda")
    outs = []
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ex = mx.CudaExecutor(s)
        o = torch.zeros_like(x)
        fn = lambda: mx.make_tensor(o).set(mx.cumsum(mx.make_tensor(x))).run(ex)  # noqa: E731
        for _ in range(3):
            fn()
            ex.sync()
            outs.append(o.clone())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
        for _ in range(3):
            o.zero_()
            g.replay()
            s.synchronize()
            outs.append(o.clone())
    for t in outs[1:]:
        assert torch.equal(t, outs[0])
    truth = torch.cumsum(x.double(), 0)
    assert float(((outs[0].double() - truth).abs() / truth).max()) <= 1e-5
