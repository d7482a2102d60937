"""-m gpu: BASELINE.json configs 2 and 1 AT FULL SIZE against THE REFERENCE ITSELF — oracle/_ref/libmatx_ref_host.so is the
unmodified MatX HostExecutor<ThreadsMode::ALL> (oracle/ref_wrap.cu, built where /root/reference exists; the library travels
to the GPU box).  The same host bits are fed to both sides.

  C2  full-tensor sum / mean / max / min / argmax / argmin / any / all of 2^30 fp32 values drawn from U[0,1): the 24-bit
      mantissa makes the extrema occur dozens of times, so argmax / argmin exercise the lowest-index rule on natural ties.
      Exact ops are compared bit for bit including the index.  Sums: the bar is 1e-5 relative; the reference accumulates
      2^30 / threads values sequentially in fp32, which is itself lossy at this size (tier 3), so the
      device result is held to 1e-5 of fp64 truth and to the reference within the reference's OWN error against that truth.
  C1  sum(a*b+c, {1}) of 16384 x 4096 fp32: every row within 1e-5 of the reference's value.
Mirrors test/00_operators/ReductionTests.cu:935-983,1241-1387 (same statements, benchmark sizes)."""
import ctypes as C
import os

import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests import oracle_harness as H

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    r = H.load_ref_host()
    if r is None:
        pytest.skip("oracle/_ref/libmatx_ref_host.so not built (needs /root/reference at build time)")
    return r


def _host_uniform(n, seed):
    from concurrent.futures import ThreadPoolExecutor
    x = np.empty(n, np.float32)
    blk = 1 << 24

    def fill(i):
        np.random.default_rng(seed * 1000 + i).random(out=x[i * blk:(i + 1) * blk], dtype=np.float32)
    with ThreadPoolExecutor(max_workers=min(16, os.cpu_count() or 1)) as ex:
        list(ex.map(fill, range((n + blk - 1) // blk)))
    return x


def test_config2_full_tensor_2pow30_against_the_live_reference(ref):
    import torch
    n = 1 << 30
    x = _host_uniform(n, 4)
    dx = torch.from_numpy(x).cuda()
    tx = mx.make_tensor(dx)
    ex = mx.CudaExecutor()
    truth = float(dx.double().sum().item())

    def ours(op):
        o = torch.zeros((), device="cuda")
        i = torch.full((), -1, dtype=torch.int64, device="cuda")
        r = getattr(mx, op)(tx)
        (mx.mtie(mx.make_tensor(o), mx.make_tensor(i)) if op.startswith("arg") else mx.make_tensor(o)).set(r).run(ex)
        ex.sync()
        return o.item(), i.item(), ex.last_kernel()

    def theirs(op):
        out, idx = np.zeros((), np.float32), np.full((), -1, np.int64)
        rc = ref.reduce(getattr(A, "RED_" + op.upper()), x.ctypes.data, A.F32, [n], [1], [0], out.ctypes.data, idx.ctypes.data, 1, mode=1)
        assert rc == 0
        return float(out), int(idx)

    for op in ("max", "min", "any", "all"):
        (g, _, k), (w, _) = ours(op), theirs(op)
        assert g == w, (op, k, g, w)
    for op in ("argmax", "argmin"):
        (g, gi, k), (w, wi) = ours(op), theirs(op)
        assert g == w and gi == wi, (op, k, (g, gi), (w, wi))
        ties = int((dx == g).sum().item())
        assert ties > 1, "expected natural ties at the extremum of 2^30 24-bit uniforms"
        assert int(torch.nonzero(dx == g)[0].item()) == gi          # and it IS the lowest index holding the value
    for op, scale in (("sum", 1.0), ("mean", 1.0 / n)):
        (g, _, k), (w, _) = ours(op), theirs(op)
        t = truth * scale
        ref_err = abs(w - t) / t
        assert abs(g - t) <= 1e-5 * t, (op, k, g, t)
        assert abs(g - w) <= (1e-5 + ref_err) * t, (op, k, g, w, ref_err)
        print("C2 %s: device %.9g, reference %.9g (its own rel err vs fp64 %.2e), fp64 %.12g" % (op, g, w, ref_err, t))


def test_config1_fma_rowsum_16384x4096_against_the_live_reference(ref):
    import torch
    rows, cols = 16384, 4096
    a, b = _host_uniform(rows * cols, 1).reshape(rows, cols), _host_uniform(rows * cols, 2).reshape(rows, cols)
    c = (_host_uniform(rows * cols, 3) - np.float32(0.5)).reshape(rows, cols)
    want = np.zeros(rows, np.float32)
    f = ref.fn("mref_fma_sum")
    f.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
    assert f(1, a.ctypes.data, b.ctypes.data, c.ctypes.data, want.ctypes.data, rows, cols) == 0
    da, db, dc = (torch.from_numpy(t).cuda() for t in (a, b, c))
    out = torch.zeros(rows, device="cuda")
    ex = mx.CudaExecutor()
    mx.make_tensor(out).set(mx.sum(mx.make_tensor(da) * mx.make_tensor(db) + mx.make_tensor(dc), [1])).run(ex)
    ex.sync()
    got = out.cpu().numpy()
    err = np.max(np.abs(got.astype(np.float64) - want) / np.abs(want))
    assert err <= 1e-5, (ex.last_kernel(), err)
    # the other row statements of the family on the same data, exact ones bit for bit
    for op in ("max", "min", "argmax", "argmin"):
        o = torch.zeros(rows, device="cuda")
        i = torch.zeros(rows, dtype=torch.int64, device="cuda")
        r = getattr(mx, op)(mx.make_tensor(da), [1])
        (mx.mtie(mx.make_tensor(o), mx.make_tensor(i)) if op.startswith("arg") else mx.make_tensor(o)).set(r).run(ex)
        ex.sync()
        w, wi = ref.reduce_np(getattr(A, "RED_" + op.upper()), a, [1], mode=1)
        assert np.array_equal(o.cpu().numpy(), w), op
        if wi is not None:
            assert np.array_equal(i.cpu().numpy(), wi), op
