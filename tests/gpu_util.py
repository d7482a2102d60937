"""Helpers for the -m gpu parity tests: state ONE MatX statement, run it through the C ABI on the device and
through the CPU oracle on the same bits, and hand both results back."""
from __future__ import annotations

import numpy as np

from matx_b200 import _abi as A
from matx_b200 import ops as mx
from tests.oracle_harness import np_tensor

_EXEC = None


def executor() -> mx.CudaExecutor:
    global _EXEC
    if _EXEC is None:
        _EXEC = mx.CudaExecutor()
    return _EXEC


_NP_OF = {A.F32: np.float32, A.F64: np.float64, A.C64: np.complex64, A.I32: np.int32, A.I64: np.int64, A.U8: np.uint8,
          A.BF16: np.uint16, A.F16: np.uint16}


def to_dev(arr: np.ndarray, dtype: int | None = None):
    """numpy -> torch cuda tensor with the same bits (bf16 / f16 travel as uint16 bit patterns)."""
    import torch
    a = np.ascontiguousarray(arr)
    if dtype == A.BF16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.bfloat16)
    if dtype == A.F16:
        return torch.from_numpy(a.view(np.int16)).cuda().view(torch.float16)
    return torch.from_numpy(a).cuda()


def from_dev(t, dtype: int | None = None) -> np.ndarray:
    import torch
    if dtype in (A.BF16, A.F16):
        return t.view(torch.int16).cpu().numpy().view(np.uint16)
    return t.cpu().numpy()


def run_reduce(oracle, build, arrays, out_dtype: int, dtypes=None, half_acc: int = -1):
    """build(*tensors) -> ReduceExpr.  Returns (got, got_idx, want, want_idx, kernel_name)."""
    import torch
    dtypes = dtypes or [None] * len(arrays)
    dev = [to_dev(a, d) for a, d in zip(arrays, dtypes)]
    r = build(*[mx.make_tensor(t) for t in dev])
    want_idx = r.op in (A.RED_ARGMAX, A.RED_ARGMIN)
    npdt = _NP_OF[out_dtype]
    tdt = {A.BF16: torch.bfloat16, A.F16: torch.float16}.get(out_dtype)
    out_d = torch.full(r.out_shape, -77, dtype=tdt, device="cuda") if tdt else torch.from_numpy(np.full(r.out_shape, 77, npdt)).cuda()
    idx_d = torch.full(r.out_shape, -7, dtype=torch.int64, device="cuda") if want_idx else None
    ex = executor()
    if want_idx:
        mx.mtie(mx.make_tensor(out_d), mx.make_tensor(idx_d)).set(r).run(ex)
    else:
        mx.make_tensor(out_d).set(r).run(ex)
    ex.sync()
    kname = ex.last_kernel()
    got = from_dev(out_d, out_dtype)
    got_idx = idx_d.cpu().numpy() if want_idx else None

    r2 = build(*[np_tensor(np.ascontiguousarray(a), d) for a, d in zip(arrays, dtypes)])
    want = np.zeros(r2.out_shape, npdt)
    widx = np.zeros(r2.out_shape, np.int64) if want_idx else None
    oracle.reduce(r2, want, widx, out_dtype=out_dtype, half_acc=half_acc)
    return got, got_idx, want, widx, kname


def run_elementwise(oracle, build, arrays, out_shape, out_dtype: int, dtypes=None):
    """build(*tensors) -> Op.  Returns (got, want, kernel_name)."""
    import torch
    dtypes = dtypes or [None] * len(arrays)
    dev = [to_dev(a, d) for a, d in zip(arrays, dtypes)]
    rhs = build(*[mx.make_tensor(t) for t in dev])
    npdt = _NP_OF[out_dtype]
    tdt = {A.BF16: torch.bfloat16, A.F16: torch.float16}.get(out_dtype)
    out_d = torch.zeros(out_shape, dtype=tdt, device="cuda") if tdt else torch.from_numpy(np.zeros(out_shape, npdt)).cuda()
    ex = executor()
    mx.make_tensor(out_d).set(rhs).run(ex)
    ex.sync()
    kname = ex.last_kernel()
    got = from_dev(out_d, out_dtype)
    rhs2 = build(*[np_tensor(np.ascontiguousarray(a), d) for a, d in zip(arrays, dtypes)])
    want = np.zeros(out_shape, npdt)
    oracle.elementwise(rhs2, want, out_dtype)
    return got, want, kname


def rel_err(got, want):
    got = np.asarray(got, dtype=np.complex128 if np.iscomplexobj(got) else np.float64)
    want = np.asarray(want, dtype=got.dtype)
    den = np.maximum(np.abs(want), 1e-30)
    return float(np.max(np.abs(got - want) / den)) if got.size else 0.0
