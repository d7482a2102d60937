"""-m gpu: find / find_idx (stream compaction) through the C ABI against the CPU oracle (orc_find, pinned to the
reference's HostExecutor by tests/test_find_oracle.py) on the same bits: every selection functor, ties with the
threshold, tile-boundary sizes, unaligned / strided / sliced / transposed views, integer and fp64 values, an expression
operand, a capacity smaller than the count, empty input and empty selection; and at 2^28 elements against torch's
masked_select (bit-exact, stable order)."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import np_tensor

pytestmark = pytest.mark.gpu

SEL = [mx.LT, mx.GT, mx.EQ, mx.NEQ, mx.LTE, mx.GTE]


def run_find(oracle, view_of, arr, sel, want_idx, cap=None, idx_dtype=np.int32):
    """view_of(tensor) -> operand.  Returns (got, n_got, want, n_want, kernel)."""
    import torch
    dev = G.to_dev(arr)
    opd = view_of(mx.make_tensor(dev))
    n_el = int(np.prod(opd.shape)) if len(opd.shape) else 1
    cap = n_el if cap is None else cap
    odt = idx_dtype if want_idx else arr.dtype
    out_d = torch.from_numpy(np.full(max(cap, 1), -7, odt)).cuda()[:cap]
    nf_d = torch.full((), -3, dtype=torch.int32, device="cuda")
    ex = G.executor()
    f = mx.find_idx if want_idx else mx.find
    mx.mtie(mx.make_tensor(out_d), mx.make_tensor(nf_d)).set(f(opd, sel)).run(ex)
    ex.sync()
    k = ex.last_kernel()
    want = np.full(max(cap, 1), -7, odt)[:cap]
    wn = oracle.find(f(view_of(np_tensor(arr)), sel), want)
    return out_d.cpu().numpy(), int(nf_d.item()), want, wn, k


def assert_same(res):
    got, n, want, wn, k = res
    assert n == wn, (k, n, wn)
    m = min(n, len(got))
    assert np.array_equal(got[:m], want[:m]), (k, got[:8], want[:8])
    assert np.all(got[m:] == -7), k      # nothing written past the selection


def test_reference_test_bodies_on_device(oracle):
    rng = np.random.default_rng(1)
    t1 = (rng.random(100) * 2).astype(np.float32)        # ReductionTests.cu:1624-1647, 1665-1688
    assert_same(run_find(oracle, lambda t: t, t1, mx.GT(0.5), False))
    res = run_find(oracle, lambda t: t, t1, mx.GT(0.5), True)
    assert_same(res)
    assert res[4].startswith("select|") and res[4].endswith("aot"), res[4]


@pytest.mark.parametrize("n", [1, 5, 4095, 4096, 4097, 16384, 100_003, (1 << 20) + 3])
@pytest.mark.parametrize("sel", range(6))
def test_sizes_and_functors(oracle, n, sel):
    rng = np.random.default_rng(100 + sel)
    x = (rng.integers(0, 9, n) * 0.25).astype(np.float32)   # ties with the threshold everywhere
    assert_same(run_find(oracle, lambda t: t, x, SEL[sel](1.0), False))
    assert_same(run_find(oracle, lambda t: t, x, SEL[sel](1.0), True, idx_dtype=np.int64 if sel % 2 else np.int32))


def test_views(oracle):
    rng = np.random.default_rng(7)
    x = (rng.integers(0, 9, (300, 257)) * 0.25).astype(np.float32)
    flat = x.reshape(-1).copy()
    for view_of in (lambda t: t,                                   # contiguous 2-D: collapses to one dim
                    lambda t: t.Slice([3, 5], [290, 250]),         # sliced: two dims, scalar walk
                    lambda t: mx.permute(t, [1, 0])):              # transposed
        for want_idx in (False, True):
            assert_same(run_find(oracle, view_of, x, mx.GTE(1.25), want_idx))
    # unaligned start and a strided 1-D view
    assert_same(run_find(oracle, lambda t: t.Slice([1], [len(flat)]), flat, mx.LT(0.75), False))
    assert_same(run_find(oracle, lambda t: t.Slice([1], [len(flat)]), flat, mx.LT(0.75), True))


def test_other_value_types_and_expressions(oracle):
    rng = np.random.default_rng(8)
    xi = rng.integers(-50, 50, 70_001).astype(np.int32)
    for sel in (mx.GT(10), mx.EQ(-3), mx.LTE(0)):
        assert_same(run_find(oracle, lambda t: t, xi, sel, False))
        assert_same(run_find(oracle, lambda t: t, xi, sel, True))
    xd = rng.standard_normal(50_000)
    assert_same(run_find(oracle, lambda t: t, xd, mx.GT(0.25), False))
    assert_same(run_find(oracle, lambda t: t, xd, mx.NEQ(0.0), True, idx_dtype=np.int64))
    xf = rng.random(30_000).astype(np.float32)
    res = run_find(oracle, lambda t: t * 2.0 - 1.0, xf, mx.GT(0.5), False)   # an expression as the operand (JIT)
    assert_same(res)
    assert res[4].endswith("jit"), res[4]


def test_capacity_empty_selection_and_empty_input(oracle):
    rng = np.random.default_rng(9)
    x = rng.random(20_000).astype(np.float32)
    got, n, want, wn, k = run_find(oracle, lambda t: t, x, mx.GT(0.5), False, cap=100)   # more found than fit: counted, not written
    assert n == wn == int((x > 0.5).sum()) and np.array_equal(got, want)
    got, n, want, wn, k = run_find(oracle, lambda t: t, x, mx.GT(2.0), True)
    assert n == 0 and np.all(got == -7)
    import torch
    ex = G.executor()
    e = torch.zeros(0, device="cuda")
    o = torch.zeros(4, device="cuda")
    nf = torch.full((), 5, dtype=torch.int32, device="cuda")
    mx.mtie(mx.make_tensor(o), mx.make_tensor(nf)).set(mx.find(mx.make_tensor(e), mx.GT(0.0))).run(ex)
    ex.sync()
    assert nf.item() == 0


def test_error_convention():
    import torch
    ex = G.executor()
    x = mx.make_tensor(torch.zeros(8, device="cuda"))
    o = mx.make_tensor(torch.zeros(8, device="cuda"))
    nf = mx.make_tensor(torch.zeros((), dtype=torch.int32, device="cuda"))
    with pytest.raises(TypeError):
        mx.mtie(o, nf).set(mx.find(x, lambda v: v > 0))                      # a callable cannot cross the C ABI
    with pytest.raises(TypeError):
        mx.mtie(o, mx.make_tensor(torch.zeros(1, dtype=torch.int32, device="cuda"))).set(mx.find(x, mx.GT(0.0)))   # rank-0 count
    with pytest.raises(A.MatxB200Error) as ei:
        mx.mtie(o, nf).set(mx.find_idx(x, mx.GT(0.0))).run(ex)               # indices need an integer output
    assert ei.value.status == A.ERR_INVALID
    c = mx.make_tensor(torch.zeros(8, dtype=torch.complex64, device="cuda"))
    with pytest.raises(A.MatxB200Error) as ei:
        mx.mtie(o, nf).set(mx.find(c, mx.GT(0.0))).run(ex)                   # complex has no order
    assert ei.value.status == A.ERR_NOT_SUPPORTED


@pytest.mark.parametrize("sel", range(6))
def test_single_pass_and_two_pass_kernels_agree_with_the_oracle(oracle, monkeypatch, sel):
    """1-D views take the single-pass look-back kernel (T3 values / T4 indices); MXB_SEL_TWO_PASS=1 forces the count +
    scatter pair that serves N-D views.  Both against the oracle, NaN included (NEQ accepts it, ordered comparisons do not)."""
    rng = np.random.default_rng(300 + sel)
    for n in (1, 4097, 3 * 4096, (1 << 20) + 3):
        x = (rng.integers(0, 9, n) * 0.25).astype(np.float32)
        x[:: 7] = np.nan if sel == 3 else x[:: 7]
        for two_pass in ("0", "1"):
            monkeypatch.setenv("MXB_SEL_TWO_PASS", two_pass)
            for want_idx in (False, True):
                got, n_got, want, wn, k = run_find(oracle, lambda t: t, x, SEL[sel](1.0), want_idx)
                assert n_got == wn and np.array_equal(got[:wn], want[:wn], equal_nan=True), k
                want_t = (2 if want_idx else 1) if two_pass == "1" else (4 if want_idx else 3)
                assert "|T%d|" % want_t in k, k


def test_single_pass_many_tiles_and_repeated_launches(oracle):
    """More tiles than resident CTAs (look-back across several probes of 32 tiles), all-selected and none-selected tiles,
    and back-to-back launches on one handle (the status words are epoch-coded: nothing is cleared in between)."""
    import torch
    ex = mx.CudaExecutor()
    n = (1 << 24) + 12345
    rng = np.random.default_rng(77)
    x = rng.random(n, dtype=np.float32)
    x[: 1 << 20] = 2.0                       # 256 tiles in a row with everything selected
    x[1 << 22: (1 << 22) + (1 << 21)] = -1.0  # 512 tiles with nothing selected
    dx = torch.from_numpy(x).cuda()
    for rep, thr in enumerate((0.5, 0.999, 0.5, 0.0, 1.5, 0.25)):
        out = torch.full((n,), -7.0, device="cuda")
        nf = torch.zeros((), dtype=torch.int32, device="cuda")
        mx.mtie(mx.make_tensor(out), mx.make_tensor(nf)).set(mx.find(mx.make_tensor(dx), mx.GT(thr))).run(ex)
        ex.sync()
        want = x[x > np.float32(thr)]
        assert "|T3|" in ex.last_kernel() and nf.item() == want.size, (rep, ex.last_kernel(), nf.item(), want.size)
        assert np.array_equal(out[: want.size].cpu().numpy(), want) and bool((out[want.size:] == -7.0).all().item())
        idx = torch.full((n,), -7, dtype=torch.int64, device="cuda")
        mx.mtie(mx.make_tensor(idx), mx.make_tensor(nf)).set(mx.find_idx(mx.make_tensor(dx), mx.LTE(thr))).run(ex)
        ex.sync()
        wi = np.nonzero(x <= np.float32(thr))[0]
        assert nf.item() == wi.size and np.array_equal(idx[: wi.size].cpu().numpy(), wi)


def test_output_capacity_is_respected_by_the_single_pass_kernel():
    import torch
    ex = mx.CudaExecutor()
    x = torch.arange(100000, device="cuda", dtype=torch.float32)
    out = torch.full((1000 + 64,), -7.0, device="cuda")
    nf = torch.zeros((), dtype=torch.int32, device="cuda")
    mx.mtie(mx.make_tensor(out[:1000]), mx.make_tensor(nf)).set(mx.find(mx.make_tensor(x), mx.GTE(10.0))).run(ex)
    ex.sync()
    assert nf.item() == 100000 - 10                                   # counted in full ...
    assert torch.equal(out[:1000], x[10:1010]) and bool((out[1000:] == -7.0).all().item())   # ... written up to the capacity


def test_full_size_against_masked_select():
    import torch
    ex = mx.CudaExecutor()
    n = 1 << 28
    g = torch.Generator(device="cuda")
    g.manual_seed(21)
    x = torch.rand(n, device="cuda", generator=g)
    for thr, frac in ((0.99, 0.01), (0.5, 0.5)):
        want = torch.masked_select(x, x > thr)
        out = torch.empty(n, device="cuda")
        idx = torch.empty(n, dtype=torch.int64, device="cuda")
        nf = torch.zeros((), dtype=torch.int32, device="cuda")
        st = mx.mtie(mx.make_tensor(out), mx.make_tensor(nf)).set(mx.find(mx.make_tensor(x), mx.GT(thr)))
        st.run(ex)
        ex.sync()
        k = ex.last_kernel()
        assert nf.item() == want.numel(), (k, nf.item(), want.numel())
        assert torch.equal(out[: want.numel()], want), k
        mx.mtie(mx.make_tensor(idx), mx.make_tensor(nf)).set(mx.find_idx(mx.make_tensor(x), mx.GT(thr))).run(ex)
        ex.sync()
        assert nf.item() == want.numel() and torch.equal(x[idx[: want.numel()]], want)
        assert bool((idx[1: want.numel()] > idx[: want.numel() - 1]).all().item())      # stable: indices strictly increase
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            st.run(ex)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print("find(x > %g) fp32 2^28: %.4f ms = %.0f GB/s of algorithmic bytes (1 read + %.0f%% written)"
              % (thr, ms, (n * 4 * (1 + frac)) / ms / 1e6, frac * 100))


# ---- warp-tile geometry (round 2): tile = 2048 elements (1024 for 8-byte values, 256 on the scalar walk), a group is 32
# tiles, a supergroup 1024; sizes on and around every edge, sparse (bit walk) and dense (staged) tiles in one launch ----
@pytest.mark.parametrize("n", [2047, 2048, 2049, 65535, 65536, 65537, (1 << 21) - 1, 1 << 21, (1 << 21) + 1, 3 * (1 << 21) + 2049])
def test_warp_tile_group_and_supergroup_edges(oracle, n):
    rng = np.random.default_rng(n)
    x = rng.random(n).astype(np.float32)
    x[: n // 3] = np.where(rng.random(n // 3) < 0.02, 2.0, 0.0)     # sparse tiles: a few selected elements each
    x[n // 3: n // 2] = 2.0                                           # dense tiles: everything selected
    for sel in (mx.GT(1.0), mx.LTE(0.5)):
        assert_same(run_find(oracle, lambda t: t, x, sel, False))
        assert_same(run_find(oracle, lambda t: t, x, sel, True, idx_dtype=np.int64))
    assert_same(run_find(oracle, lambda t: t, x, mx.GT(1.0), True, cap=max(1, n // 5)))     # capacity below the count


@pytest.mark.parametrize("npdt", [np.float64, np.int64, np.uint8, np.int32])
def test_warp_tiles_of_other_value_widths(oracle, npdt):
    """8-byte values take 1024-element tiles (two-element vectors), uint8 sixteen-element vectors with 16-bit packed counts."""
    rng = np.random.default_rng(17)
    for n in (1023, 1024, 1025, 70_003, (1 << 20) + 5):
        x = rng.integers(0, 7, n).astype(npdt)
        for sel in (mx.GT(3), mx.NEQ(2)):
            if npdt is not np.uint8:      # the harness pre-fills outputs with -7: indices only for the unsigned type
                assert_same(run_find(oracle, lambda t: t, x, sel, False))
            assert_same(run_find(oracle, lambda t: t, x, sel, True))


def test_scalar_walk_tiles_on_strided_operands(oracle):
    """a 1-D view with a stride takes the single-pass kernel on its scalar walk (256-element warp tiles)."""
    rng = np.random.default_rng(18)
    flat = (rng.integers(0, 9, 300_001) * 0.25).astype(np.float32)

    def every_third(t):
        return mx.Tensor(t.data_ptr, t.dtype, [100_000], [3], t._keep)

    for want_idx in (False, True):
        res = run_find(oracle, every_third, flat, mx.LT(1.0), want_idx)
        assert_same(res)
        assert "|V1|" in res[4] and ("|T3|" in res[4] or "|T4|" in res[4]), res[4]
    m = (rng.integers(0, 9, (4100, 6)) * 0.25).astype(np.float32)

    def every_other_column(t):   # collapses to ONE strided dim of 4100 * 3 elements
        return mx.Tensor(t.data_ptr, t.dtype, [4100, 3], [6, 2], t._keep)

    assert_same(run_find(oracle, every_other_column, m, mx.GTE(1.0), True))
