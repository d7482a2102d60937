"""-m gpu: sort / unique / hist through the C ABI against the CPU oracle on the same bits —
bit-exact (keys are only moved; counts are integers).  Mirrors test/00_tensor/CUBTests.cu:153-260,1140-1160 and
test/00_operators/ReductionTests.cu:1744-1770."""
import zlib

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import np_tensor

pytestmark = pytest.mark.gpu
NPDT = {A.F32: np.float32, A.F64: np.float64, A.I32: np.int32, A.I64: np.int64}


def keys(rng, shape, dt, ties):
    if dt in (A.I32, A.I64):
        return rng.integers(-40 if ties else -2**31 + 1, 40 if ties else 2**31 - 1, shape).astype(NPDT[dt])
    if ties:
        return (rng.integers(-30, 30, shape) * 0.25).astype(NPDT[dt])
    x = rng.standard_normal(shape) * 10.0 ** rng.integers(-3, 6, shape)
    return x.astype(NPDT[dt])


def run_sort(oracle, x, direction, view=lambda t: t):
    import torch
    dx = G.to_dev(x)
    opd = view(mx.make_tensor(dx))
    out = torch.from_numpy(np.zeros(opd.shape, x.dtype)).cuda()
    ex = G.executor()
    mx.make_tensor(out).set(mx.sort(opd, direction)).run(ex)
    ex.sync()
    k = ex.last_kernel()
    want = np.zeros(opd.shape, x.dtype)
    oracle.sort(mx.sort(view(np_tensor(x)), direction), want)
    return out.cpu().numpy(), want, k


@pytest.mark.parametrize("shape", [(1,), (2,), (33,), (1000,), (4096,), (4097,), (70001,), (1 << 20,), (37, 129), (5, 5000), (300, 64), (3, 40000), (4, 6, 200)])
@pytest.mark.parametrize("dt", [A.F32, A.I32, A.F64, A.I64])
def test_sort_rows_bit_exact(oracle, dt, shape):
    rng = np.random.default_rng(zlib.crc32(repr((dt, shape)).encode()))
    for ties in (False, True):
        x = keys(rng, shape, dt, ties)
        for direction in (mx.SORT_DIR_ASC, mx.SORT_DIR_DESC):
            got, want, k = run_sort(oracle, x, direction)
            assert np.array_equal(got, want), (k, shape, ties, direction)
            assert k.startswith("sort_bitonic" if shape[-1] <= 4096 else "sort_radix"), k


def test_sort_special_values_and_views(oracle):
    x = np.array([np.inf, -np.inf, 3.0, -3.0, 1e-38, -1e-38, np.finfo(np.float32).max, -np.finfo(np.float32).max, 2.5, 2.5, 0.0, 1.0] * 700, np.float32)
    for direction in (mx.SORT_DIR_ASC, mx.SORT_DIR_DESC):
        got, want, k = run_sort(oracle, x, direction)
        assert np.array_equal(got, want), k
    rng = np.random.default_rng(5)
    a = rng.standard_normal((64, 9000)).astype(np.float32)
    got, want, k = run_sort(oracle, a, mx.SORT_DIR_ASC, view=lambda t: t * 2.0 - 1.0)      # an expression is evaluated into the output first
    assert np.array_equal(got, want), k
    got, want, k = run_sort(oracle, a, mx.SORT_DIR_ASC, view=lambda t: mx.permute(t, [1, 0]))   # a permuted view: rows of the view
    assert np.array_equal(got, want), k


def test_sort_full_size_property():
    """2^26 fp32 keys: sortedness and multiset preservation (sum of a checksum of the bits), both directions."""
    import torch
    ex = mx.CudaExecutor()
    n = 1 << 26
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    x = torch.randn(n, device="cuda", generator=g)
    out = torch.empty_like(x)
    for direction in (mx.SORT_DIR_ASC, mx.SORT_DIR_DESC):
        mx.make_tensor(out).set(mx.sort(mx.make_tensor(x), direction)).run(ex)
        ex.sync()
        d = out[1:] - out[:-1]
        assert bool((d >= 0).all().item()) if direction == mx.SORT_DIR_ASC else bool((d <= 0).all().item())
        assert out.view(torch.int32).to(torch.int64).sum().item() == x.view(torch.int32).to(torch.int64).sum().item()
        assert torch.equal(out, torch.sort(x, descending=direction == mx.SORT_DIR_DESC).values)


@pytest.mark.parametrize("n", [1, 100, 4097, 300000])
@pytest.mark.parametrize("dt", [A.F32, A.I32, A.F64])
def test_unique_matches_oracle(oracle, dt, n):
    import torch
    rng = np.random.default_rng(n + dt)
    x = keys(rng, (n,), dt, ties=True)
    dx = G.to_dev(x)
    out = torch.from_numpy(np.full(n, -7, x.dtype)).cuda()
    nf = torch.zeros((), dtype=torch.int32, device="cuda")
    ex = G.executor()
    mx.mtie(mx.make_tensor(out), mx.make_tensor(nf)).set(mx.unique(mx.make_tensor(dx))).run(ex)
    ex.sync()
    want = np.full(n, -7, x.dtype)
    wn = oracle.unique(mx.unique(np_tensor(x)), want)
    assert nf.item() == wn and np.array_equal(out.cpu().numpy()[:wn], want[:wn]), ex.last_kernel()
    assert np.array_equal(want[:wn], np.unique(x))
    assert bool((out[wn:] == -7).all().item())          # nothing written past the distinct values


def test_unique_known_answer_of_the_reference():
    import torch
    ex = G.executor()
    t = torch.arange(100, device="cuda", dtype=torch.float32) % 10          # ReductionTests.cu:1744-1770
    out = torch.zeros(100, device="cuda")
    nf = torch.zeros((), dtype=torch.int32, device="cuda")
    mx.mtie(mx.make_tensor(out), mx.make_tensor(nf)).set(mx.unique(mx.make_tensor(t))).run(ex)
    ex.sync()
    assert nf.item() == 10 and out[:10].tolist() == list(range(10))


def run_hist(oracle, x, lower, upper, levels, view=lambda t: t):
    import torch
    dx = G.to_dev(x)
    r = mx.hist(view(mx.make_tensor(dx)), lower, upper, levels)
    out = torch.full(r.out_shape, -5, dtype=torch.int32, device="cuda")
    ex = G.executor()
    mx.make_tensor(out).set(r).run(ex)
    ex.sync()
    want = np.full(r.out_shape, -5, np.int32)
    oracle.hist(mx.hist(view(np_tensor(x)), lower, upper, levels), want)
    return out.cpu().numpy(), want, ex.last_kernel()


def test_hist_known_answers_of_the_reference(oracle):
    x = np.array([2.2, 6.0, 7.1, 2.9, 3.5, 0.3, 2.9, 2.0, 6.1, 999.5], np.float32)       # CUBTests.cu:153-176
    got, want, k = run_hist(oracle, x, 0.0, 12.0, 7)
    assert got.tolist() == [1, 5, 0, 3, 0, 0] and np.array_equal(got, want) and k.startswith("hist|"), k
    s = np.array([0, 99, 1, 99, 2, 99, 0, 99, 1, 99, 2, 99], np.float32)                    # CUBTests.cu:178-197
    import torch
    ds = torch.from_numpy(s).cuda()
    out = torch.full((3,), -5, dtype=torch.int32, device="cuda")
    ex = G.executor()
    mx.make_tensor(out).set(mx.hist(mx.make_tensor(ds[::2]), 0.0, 3.0, 4)).run(ex)       # every other element: a strided view
    ex.sync()
    assert out.tolist() == [2, 2, 2], ex.last_kernel()


@pytest.mark.parametrize("shape,levels", [((1000,), 11), ((1 << 20,), 257), ((1 << 22,), 4097), ((37, 5000), 33), ((4, 6, 3000), 8), ((100000,), 40001)])
@pytest.mark.parametrize("dt", [A.F32, A.F64, A.I32])
def test_hist_matches_oracle(oracle, dt, shape, levels):
    rng = np.random.default_rng(zlib.crc32(repr((dt, shape, levels)).encode()))
    if dt == A.I32:
        x = rng.integers(-100, 1100, shape).astype(np.int32)
    else:
        x = (rng.random(shape) * 1200 - 100).astype(NPDT[dt])
        x.ravel()[:7] = [0.0, 1000.0, np.nextafter(NPDT[dt](1000.0), NPDT[dt](0)), -0.0, 999.999, 500.0, 250.0][: min(7, x.size)]
    got, want, k = run_hist(oracle, x, 0, 1000, levels)
    assert np.array_equal(got, want), (k, int(np.abs(got - want).sum()))
    # every in-range sample is counted exactly once — except one a hair below `upper` whose (x - lower) * scale rounds
    # up to `bins` (999.99994f * 0.256f == 256.0f): CUB's formula indexes past its last bin there, here it is dropped
    inr = (x >= 0) & (x < 1000)
    if dt != A.I32:
        bins = levels - 1
        scale = NPDT[dt](bins) / NPDT[dt](1000)
        inr &= ((x - NPDT[dt](0)) * scale).astype(np.int64) < bins
    assert got.sum() == int(inr.sum())
    got2, want2, k2 = run_hist(oracle, x, 0, 1000, levels, view=lambda t: t * 1)      # fused expression operand
    assert np.array_equal(got2, want), k2
