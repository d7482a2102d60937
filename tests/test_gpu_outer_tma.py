"""-m gpu: the TMA-staged flavour of reduce_outer (`red_outer_tma`): a strided / permuted reduce dim of a plain tensor,
tiles of (reduce rows x a strip of the unit-stride batch dim) staged in a shared-memory ring by `cp.async.bulk`.
Parity against the CPU oracle at small sizes (both copy shapes: one copy per stage when the strip covers whole
contiguous rows, one copy per row otherwise; ragged last strip, ragged last chunk), bit-equality with the LDG walker
on the exact ops, and config 5 at full size against fp32 device truth over the WHOLE output."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.oracle_harness import f32_to_bf16_bits, bf16_bits_to_f32
from tests.test_gpu_parity import check, data

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _tma_on(monkeypatch):
    monkeypatch.setenv("MXB_OUTER_TMA", "1")


def test_bf16_permuted_contiguous_strips(oracle):
    # config 5 in miniature: t[j][k][i] -> out[i][j] = sum_k; the strip covers the whole 512-byte row: one copy per stage
    rng = np.random.default_rng(50)
    f = (rng.random((160, 100, 256)) * 0.25).astype(np.float32)
    bits = f32_to_bf16_bits(f).reshape(f.shape)
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.sum(mx.permute(t, [2, 0, 1]), [2]), [bits], A.BF16, dtypes=[A.BF16])
    assert k.startswith("red_outer_tma") and "|V8" in k and k.endswith("aot"), k
    g = bf16_bits_to_f32(got)
    truth = bf16_bits_to_f32(bits).astype(np.float64).sum(axis=1).T
    assert g.shape == (256, 160)
    assert np.max(np.abs(g - truth) / truth) <= 2 ** -8
    assert G.rel_err(g, bf16_bits_to_f32(want)) <= 1e-2


@pytest.mark.parametrize("shape", [(150, 70, 320), (149, 129, 1056), (300, 64, 36)])
def test_f32_permuted_all_ops(oracle, shape):
    # 320 columns = 80 chunks in a 128-chunk strip (idle threads); 1056 = 264 chunks: strips of 128 / 128 / 8 or narrower;
    # 36 columns = 9 chunks: too narrow for the ring -> the LDG walker keeps the shape
    rng = np.random.default_rng(51 + shape[2])
    x = data(rng, shape, A.F32, ties=True)
    for op in ["max", "min", "argmax", "argmin", "any", "all"]:
        k = check(oracle, op, lambda t, op=op: getattr(mx, op)(mx.permute(t, [2, 0, 1]), [2]), [x], A.F32)
        assert k.startswith("red_outer_tma") == (shape[2] >= 128), k
    y = data(rng, shape, A.F32) + np.float32(0.5)
    for op in ["sum", "mean"]:
        check(oracle, op, lambda t, op=op: getattr(mx, op)(mx.permute(t, [2, 0, 1]), [2]), [y], A.F32, tol=2e-5)


def test_column_reductions_row_copies(oracle):
    # sum(x, {0}) of [130, 20000]: 5000 chunks per row -> strips of 32 chunks (157 items), pitch 80000 B != strip
    # bytes: one bulk copy per row; last strip has 8 chunks; 130 rows = chunks of 64 + 64 + 2
    rng = np.random.default_rng(52)
    x = data(rng, (130, 20000), A.F32, ties=True)
    for op in ["max", "argmax", "argmin", "all"]:
        k = check(oracle, op, lambda t, op=op: getattr(mx, op)(t, [0]), [x], A.F32)
        assert k.startswith("red_outer_tma"), k
    y = data(rng, (130, 20000), A.F32) + np.float32(0.5)
    got, _, want, _, k = G.run_reduce(oracle, lambda t: mx.sum(t, [0]), [y], A.F32)
    truth = y.astype(np.float64).sum(0)
    assert k.startswith("red_outer_tma"), k
    assert np.max(np.abs(got - truth) / truth) <= 1e-5, k


def test_sliced_view_and_int32(oracle):
    # a slice keeps 16-byte alignment (offsets multiple of 4 fp32): rows of the view are NOT contiguous in memory
    rng = np.random.default_rng(53)
    x = data(rng, (96, 24000), A.F32, ties=True)
    k = check(oracle, "argmax", lambda t: mx.argmax(t.Slice([8, 64], [88, 23488]), [0]), [x], A.F32)
    assert k.startswith("red_outer_tma"), k
    k = check(oracle, "min", lambda t: mx.min(t.Slice([0, 4], [96, 23996]), [0]), [x], A.F32)
    assert k.startswith("red_outer_tma"), k
    xi = data(rng, (100, 19200), A.I32)
    for op in ["sum", "max", "argmin"]:
        k = check(oracle, op, lambda t, op=op: getattr(mx, op)(t, [0]), [xi], A.I32, tol=1e-30)   # integer sums are exact
    assert k.startswith("red_outer_tma"), k


@pytest.mark.parametrize("mode", ["0", "1"])
def test_both_copy_modes(oracle, monkeypatch, mode):
    """MXB_OUTER_TMA_MODE=1: one tensor-map tile copy per stage (UTMALDG); =0: cp.async.bulk copies, one per stage when
    the strip is whole contiguous rows, else one per row.  Ragged strips and ragged last chunks in both."""
    monkeypatch.setenv("MXB_OUTER_TMA_MODE", mode)
    rng = np.random.default_rng(60)
    x = data(rng, (151, 75, 1056), A.F32, ties=True)          # 264 chunks per row: strips of 32, the last one 8 wide
    for op in ["argmax", "min", "all"]:
        k = check(oracle, op, lambda t, op=op: getattr(mx, op)(mx.permute(t, [2, 0, 1]), [2]), [x], A.F32)
        assert k.startswith("red_outer_tma"), k
    y = data(rng, (200, 67, 512), A.F32, ties=True)           # whole 2 KB rows per strip: contiguous
    for op in ["argmin", "max"]:
        k = check(oracle, op, lambda t, op=op: getattr(mx, op)(mx.permute(t, [2, 0, 1]), [2]), [y], A.F32)
        assert k.startswith("red_outer_tma"), k
    z = data(rng, (130, 20000), A.F32, ties=True)             # no other batch dim: column reductions
    k = check(oracle, "argmax", lambda t: mx.argmax(t, [0]), [z], A.F32)
    assert k.startswith("red_outer_tma"), k


def test_same_bits_as_the_ldg_walker(oracle, monkeypatch):
    import torch
    rng = np.random.default_rng(54)
    x = G.to_dev(data(rng, (200, 96, 512), A.F32, ties=True))
    ex = G.executor()
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("MXB_OUTER_TMA", flag)
        v = torch.zeros((512, 200), device="cuda")
        i = torch.zeros((512, 200), dtype=torch.int64, device="cuda")
        mx.mtie(mx.make_tensor(v), mx.make_tensor(i)).set(mx.argmax(mx.permute(mx.make_tensor(x), [2, 0, 1]), [2])).run(ex)
        ex.sync()
        res[flag] = (v.cpu().numpy(), i.cpu().numpy(), ex.last_kernel())
    assert res["1"][2].startswith("red_outer_tma") and res["0"][2].startswith("red_outer|"), (res["1"][2], res["0"][2])
    assert np.array_equal(res["1"][0], res["0"][0]) and np.array_equal(res["1"][1], res["0"][1])


def test_config5_full_size_whole_output(monkeypatch):
    import torch
    ex = mx.CudaExecutor()
    g = torch.Generator(device="cuda")
    g.manual_seed(11)
    t = (torch.rand((1024, 1024, 1024), device="cuda", generator=g) * 0.25).to(torch.bfloat16)
    truth = torch.empty((1024, 1024), device="cuda", dtype=torch.float64)
    for j0 in range(0, 1024, 64):    # out[i][j] = sum_k t[j][k][i]
        truth[:, j0:j0 + 64] = t[j0:j0 + 64].double().sum(dim=1).T
    out = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("MXB_OUTER_TMA", flag)
        o = torch.zeros((1024, 1024), device="cuda", dtype=torch.bfloat16)
        st = mx.make_tensor(o).set(mx.sum(mx.permute(mx.make_tensor(t), [2, 0, 1]), [2]))
        st.run(ex)
        ex.sync()
        k = ex.last_kernel()
        assert k.startswith("red_outer_tma" if flag == "1" else "red_outer|"), k
        err = ((o.double() - truth).abs() / truth).max().item()
        assert err <= 2 ** -8, (k, err)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            st.run(ex)
        e0.record()
        for _ in range(10):
            st.run(ex)
        e1.record()
        torch.cuda.synchronize()
        out[flag] = (o, e0.elapsed_time(e1) / 10, k)
    print("config 5: red_outer_tma %.4f ms, red_outer (LDG) %.4f ms" % (out["1"][1], out["0"][1]))
    # fp32 accumulation on both sides, one rounding: at most one bf16 ulp apart (different summation order)
    d = (out["1"][0].float() - out["0"][0].float()).abs() / out["0"][0].float()
    assert d.max().item() <= 2 ** -7
