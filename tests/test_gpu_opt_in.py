"""-m gpu, and only with MXB_TEST_OPT_IN=1: kernels that are written and compile for sm_100a but have not been run on a
GPU yet, so they are opt-in in the library (an environment knob each) and opt-in here.  Once one has passed, its knob
becomes the default and its test moves to the regular files.
  * MXB_VAR_TMA2=1 — var_inner_tma2: producer warp + two consumer teams (twin of var_tma, same arithmetic)
  * MXB_SEL_FAST=1 — see tests/test_gpu_find.py::test_fast_instances_opt_in"""
import os

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import gpu_util as G
from tests.test_gpu_parity import check, data

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("MXB_TEST_OPT_IN") != "1", reason="opt-in kernels, not yet validated on a GPU")]


@pytest.mark.parametrize("cols,dt", [(8192, A.C64), (4096, A.C64), (2048, A.C64), (16384, A.F32), (4096, A.F32), (12000, A.F32)])
def test_var_tma2_matches_oracle(oracle, monkeypatch, cols, dt):
    monkeypatch.setenv("MXB_VAR_TMA2", "1")
    rng = np.random.default_rng(cols)
    x = data(rng, (37, cols), dt) if dt == A.C64 else (rng.random((37, cols)) + 0.5).astype(np.float32)
    for op, ddof in (("var", 1), ("stdd", 0)):
        k = check(oracle, op, lambda t, op=op, ddof=ddof: getattr(mx, op)(t, [1], ddof), [x], A.F32, tol=2e-5)
        assert k.startswith("var_tma2|"), k
    monkeypatch.setenv("MXB_VAR_TMA2", "0")
    got0 = G.run_reduce(oracle, lambda t: mx.var(t, [1], 1), [x], A.F32)[0]
    monkeypatch.setenv("MXB_VAR_TMA2", "1")
    got1 = G.run_reduce(oracle, lambda t: mx.var(t, [1], 1), [x], A.F32)[0]
    assert G.rel_err(got1, got0) <= 1e-6          # same two-pass arithmetic, a different partial-sum tree


def test_var_tma2_config3_timing(monkeypatch):
    import torch
    ex = mx.CudaExecutor()
    x = torch.view_as_complex(torch.randn(65536, 8192, 2, device="cuda"))
    out = torch.empty(65536, device="cuda")
    res = {}
    for flag in ("0", "1"):
        monkeypatch.setenv("MXB_VAR_TMA2", flag)
        st = mx.make_tensor(out).set(mx.var(mx.make_tensor(x), [1], 1))
        for _ in range(3):
            st.run(ex)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            st.run(ex)
        e1.record()
        torch.cuda.synchronize()
        res[flag] = (e0.elapsed_time(e1) / 10, ex.last_kernel(), out[:64].clone())
    print("C3 var: var_tma %.4f ms, var_tma2 %.4f ms" % (res["0"][0], res["1"][0]))
    assert res["1"][1].startswith("var_tma2|") and torch.allclose(res["0"][2], res["1"][2], rtol=1e-6)
