"""-m gpu: the fused NVLink exchange (mxb_reduce_partial_push + mxb_exchange_finalize) with `world` ranks SIMULATED on
one device: every rank's buffers live on cuda:0, so the peer stores are ordinary stores and the protocol (slots by
step parity, arrival counters, epoch, rank-order fold, lowest GLOBAL index) is exercised without a second GPU."""
import numpy as np
import pytest

from matx_b200 import _abi as A
from matx_b200 import dist as mxd
from matx_b200 import ops as mx

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_fused_exchange_simulated_ranks(world):
    import torch
    n = 300_000
    rng = np.random.default_rng(world)
    ex = mx.CudaExecutor()
    bufs = [torch.zeros(mxd.PeerExchange.buffer_bytes(world), dtype=torch.uint8, device="cuda") for _ in range(world)]
    pes = [mxd.PeerExchange(ex, world, r, _sim_buffers=bufs) for r in range(world)]
    for step in range(4):  # several steps: slot parity, counters and epochs keep working
        x = rng.integers(0, 60, n).astype(np.float32)           # many ties at the maximum
        x[rng.integers(0, n, 3)] = 60.0
        dx = torch.from_numpy(x).cuda()
        outs, plans = [], []
        for r in range(world):
            start, count = mxd.slab(n, r, world, align=64)
            t = mx.make_tensor(dx[start:start + count])
            o = [torch.zeros((), device="cuda") for _ in range(4)] + [torch.zeros((), dtype=torch.int64, device="cuda")]
            o += [torch.zeros((), device="cuda") for _ in range(2)]
            items = [(A.RED_SUM, o[0], None), (A.RED_MAX, o[1], None), (A.RED_ARGMAX, o[2], o[4]), (A.RED_MEAN, o[3], None),
                     (A.RED_VAR, o[5], None), (A.RED_STDD, o[6], None)]
            outs.append(o)
            plans.append((pes[r], pes[r].prepare(items, t, start, n)))
        # all ranks push, then all ranks fold (a real run interleaves freely; the counters make any order safe)
        A_lib = A.lib
        for pe, plan in plans:
            for op, e, off, k in plan["push"]:
                A.check(A_lib.mxb_reduce_partial_push(ex.handle, op, __import__("ctypes").byref(e), off, __import__("ctypes").byref(pe.peers), k, plan["n"]))
        for pe, plan in plans:
            A.check(A_lib.mxb_exchange_finalize(ex.handle, __import__("ctypes").byref(pe.peers), plan["fold"], plan["n"], plan["count"]))
        ex.sync()
        truth = x.astype(np.float64).sum()
        for o in outs:
            assert abs(o[0].item() - truth) <= 1e-5 * truth
            assert o[1].item() == x.max() and o[2].item() == x.max()
            assert o[4].item() == int(np.argmax(x))             # lowest GLOBAL index among ties
            assert abs(o[3].item() - truth / n) <= 1e-5 * truth / n
            v = x.astype(np.float64).var(ddof=1)
            assert abs(o[5].item() - v) <= 1e-5 * v and abs(o[6].item() - np.sqrt(v)) <= 1e-5 * np.sqrt(v)
        assert len({tuple(v.item() for v in o) for o in outs}) == 1   # every rank folded the same records in the same order


@pytest.mark.parametrize("world", [1, 4])
def test_partial_records_and_rank_order_fold(world):
    """The collective-based form: mxb_reduce_partial per slab -> (here: one buffer instead of an all-gather) -> mxb_reduce_finalize."""
    import ctypes as C
    import torch
    n = 200_000
    rng = np.random.default_rng(10 + world)
    ex = mx.CudaExecutor()
    xr = (rng.standard_normal(n) * 3 + 7).astype(np.float32)
    xc = (rng.standard_normal(n) + 1j * rng.standard_normal(n) + (2 - 1j)).astype(np.complex64)
    for x, vdt, ops in ((xr, A.F32, [A.RED_SUM, A.RED_MEAN, A.RED_MIN, A.RED_ARGMIN, A.RED_ANY, A.RED_VAR, A.RED_STDD]),
                        (xc, A.C64, [A.RED_SUM, A.RED_MEAN, A.RED_VAR])):
        dx = torch.from_numpy(x).cuda()
        for op in ops:
            gathered = torch.zeros(world * 32, dtype=torch.uint8, device="cuda")
            for r in range(world):
                start, count = mxd.slab(n, r, world, align=64)
                e = mx.lower_reduce(mx.ReduceExpr(op, mx.make_tensor(dx[start:start + count]), None))
                A.check(A.lib.mxb_reduce_partial(ex.handle, op, C.byref(e), start, C.c_void_p(gathered.data_ptr() + 32 * r)))
            real_out = op in (A.RED_VAR, A.RED_STDD, A.RED_ANY) or vdt == A.F32
            out = torch.zeros((), dtype=torch.float32 if real_out and not (vdt == A.C64 and op in (A.RED_SUM, A.RED_MEAN)) else torch.complex64, device="cuda")
            idx = torch.zeros((), dtype=torch.int64, device="cuda")
            o = mx._out_desc(mx.make_tensor(out))
            io = mx._out_desc(mx.make_tensor(idx))
            A.check(A.lib.mxb_reduce_finalize(ex.handle, op, vdt, C.c_void_p(gathered.data_ptr()), world, 32, n, 1, C.byref(o),
                                              C.byref(io) if op == A.RED_ARGMIN else None))
            ex.sync()
            x64 = x.astype(np.complex128 if vdt == A.C64 else np.float64)
            want = {A.RED_SUM: x64.sum(), A.RED_MEAN: x64.mean(), A.RED_MIN: x.min() if vdt == A.F32 else 0, A.RED_ARGMIN: x.min() if vdt == A.F32 else 0,
                    A.RED_ANY: 1.0, A.RED_VAR: x64.var(ddof=1), A.RED_STDD: x64.std(ddof=1)}[op]
            got = out.item()
            assert abs(got - want) <= 2e-5 * abs(want), (op, vdt, got, want)
            if op == A.RED_ARGMIN:
                assert idx.item() == int(np.argmin(x))


def test_steps_of_different_sizes_share_one_exchange_and_timeouts_are_reported():
    """the arrival counters are cumulative, so the fold must count what it has CONSUMED per source rank — a
    3-statement step followed by a 2-statement step (and back) on one peer table; and a step whose peer never delivers
    must not fold the stale slot: NaN / -1 in the outputs and MXB_ERR_CUDA from mxb_exchange_check."""
    import ctypes as C
    import torch
    world, n = 2, 100_000
    rng = np.random.default_rng(5)
    ex = mx.CudaExecutor()
    bufs = [torch.zeros(mxd.PeerExchange.buffer_bytes(world), dtype=torch.uint8, device="cuda") for _ in range(world)]
    pes = [mxd.PeerExchange(ex, world, r, _sim_buffers=bufs) for r in range(world)]
    for step, ops in enumerate(([A.RED_SUM, A.RED_MAX, A.RED_ARGMAX], [A.RED_MIN, A.RED_SUM], [A.RED_ARGMIN, A.RED_MAX, A.RED_SUM], [A.RED_MAX])):
        x = rng.integers(0, 50, n).astype(np.float32)
        dx = torch.from_numpy(x).cuda()
        outs, plans = [], []
        for r in range(world):
            start, count = mxd.slab(n, r, world, align=64)
            t = mx.make_tensor(dx[start:start + count])
            o = [(torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")) for _ in ops]
            items = [(op, v, i if op in (A.RED_ARGMAX, A.RED_ARGMIN) else None) for op, (v, i) in zip(ops, o)]
            outs.append(o)
            plans.append((pes[r], pes[r].prepare(items, t, start, n)))
        for pe, plan in plans:
            for op, e, off, k in plan["push"]:
                A.check(A.lib.mxb_reduce_partial_push(ex.handle, op, C.byref(e), off, C.byref(pe.peers), k, plan["n"]))
        for pe, plan in plans:
            A.check(A.lib.mxb_exchange_finalize(ex.handle, C.byref(pe.peers), plan["fold"], plan["n"], plan["count"]))
        for pe in pes:
            pe.check()
        want = {A.RED_SUM: x.astype(np.float64).sum(), A.RED_MAX: x.max(), A.RED_MIN: x.min(), A.RED_ARGMAX: x.max(), A.RED_ARGMIN: x.min()}
        for o in outs:
            for op, (v, i) in zip(ops, o):
                assert abs(v.item() - want[op]) <= 1e-5 * max(1.0, abs(want[op])), (step, op)
                if op == A.RED_ARGMAX:
                    assert i.item() == int(np.argmax(x)), step
                if op == A.RED_ARGMIN:
                    assert i.item() == int(np.argmin(x)), step


def test_missing_peer_is_an_error_not_a_stale_fold():   # waits out the 5 s peer timeout once
    import ctypes as C
    import torch
    world, n = 2, 10_000
    ex = mx.CudaExecutor()
    bufs = [torch.zeros(mxd.PeerExchange.buffer_bytes(world), dtype=torch.uint8, device="cuda") for _ in range(world)]
    pes = [mxd.PeerExchange(ex, world, r, _sim_buffers=bufs) for r in range(world)]
    dx = torch.ones(n, device="cuda")
    v, i = torch.zeros((), device="cuda"), torch.zeros((), dtype=torch.int64, device="cuda")
    plan = pes[0].prepare([(A.RED_ARGMAX, v, i)], mx.make_tensor(dx[: n // 2]), 0, n)
    for op, e, off, k in plan["push"]:     # rank 0 pushes, rank 1 never does
        A.check(A.lib.mxb_reduce_partial_push(ex.handle, op, C.byref(e), off, C.byref(pes[0].peers), k, plan["n"]))
    A.check(A.lib.mxb_exchange_finalize(ex.handle, C.byref(pes[0].peers), plan["fold"], plan["n"], plan["count"]))
    with pytest.raises(A.MatxB200Error) as ei:
        pes[0].check()
    assert ei.value.status == A.ERR_CUDA and "did not arrive" in str(ei.value)
    assert np.isnan(v.item()) and i.item() == -1
