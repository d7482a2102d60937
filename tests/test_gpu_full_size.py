"""-m gpu: the five BASELINE.json configurations AT FULL SIZE.  The CPU oracle cannot finish these in seconds, so the
checks are size-independent properties and fp64 device truth on slices: sums against fp64, extrema against a second
independent reduction, arg ops by "value at index equals the max and nothing earlier is equal" (lowest-index rule),
variance against the fp64 two-pass formula on sampled rows, linearity of sum, idempotence of max."""
import pytest

from matx_b200 import bench_configs as bc
from matx_b200 import ops as mx

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ex():
    return mx.CudaExecutor()


def test_config2_full_tensor_sum_max_argmax_2pow30(ex):
    import torch
    n = 1 << 30
    x = torch.rand(n, device="cuda")
    x[123_456_789] = 2.0            # planted unique maximum ...
    x[987_654_321] = 2.0            # ... twice: the lower index must win
    tx = mx.make_tensor(x)
    s, m, v = (torch.zeros((), device="cuda") for _ in range(3))
    i = torch.zeros((), dtype=torch.int64, device="cuda")
    mx.make_tensor(s).set(mx.sum(tx)).run(ex)
    mx.make_tensor(m).set(mx.max(tx)).run(ex)
    mx.mtie(mx.make_tensor(v), mx.make_tensor(i)).set(mx.argmax(tx)).run(ex)
    ex.sync()
    truth = x.double().sum().item()
    assert abs(s.item() - truth) <= 1e-5 * truth
    assert m.item() == 2.0 and v.item() == 2.0 and i.item() == 123_456_789
    # linearity: sum(2x) == 2 sum(x) exactly in binary floating point (scaling by 2 is exact, same reduction order)
    s2 = torch.zeros((), device="cuda")
    mx.make_tensor(s2).set(mx.sum(tx * 2.0)).run(ex)
    ex.sync()
    assert s2.item() == 2.0 * s.item()
    # argmin with ties at zero: lowest index
    x[5_000] = -1.0
    x[4_000_000_0] = -1.0
    mx.mtie(mx.make_tensor(v), mx.make_tensor(i)).set(mx.argmin(tx)).run(ex)
    ex.sync()
    assert v.item() == -1.0 and i.item() == 5_000
    # any / all over 2^30 elements
    a = torch.zeros((), device="cuda")
    mx.make_tensor(a).set(mx.all(tx)).run(ex)
    ex.sync()
    assert a.item() == float(bool((x != 0).all().item()))
    mx.make_tensor(a).set(mx.any(tx > 1.5)).run(ex)
    ex.sync()
    assert a.item() == 1.0


def test_config1_fused_fma_sum_full(ex):
    r = next(iter(bc.run_c1(ex, 6456.8).values()))
    assert r["max_rel_err_vs_fp64"] <= 1e-5 and r["kernel"].startswith("red_inner"), r


def test_config3_complex_rows_full(ex):
    res = bc.run_c3(ex, 6456.8)
    mean, var, amax = (res[k] for k in sorted(res, key=lambda k: ("mean" in k, "var" in k), reverse=True))
    assert [v for k, v in res.items() if "mean" in k][0]["max_abs_err_rows0_63"] <= 1e-6
    assert [v for k, v in res.items() if "var" in k][0]["max_rel_err_rows0_63"] <= 1e-5
    assert [v for k, v in res.items() if "argmax" in k][0]["index_match_rows0_63"] is True


def test_config4_black_scholes_full(ex):
    r = next(iter(bc.run_c4(ex, 6456.8).values()))
    assert r["max_abs_err_first_2^20_vs_fp64"] <= 1e-4 and r["kernel"].startswith("ew|"), r


def test_config5_bf16_permuted_full(ex):
    r = next(iter(bc.run_c5(ex, 6456.8).values()))
    assert r["max_rel_err_j0_7_vs_fp32"] <= 2 ** -8 + 1e-6 and r["kernel"].startswith("red_outer"), r
