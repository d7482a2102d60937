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
            items = [(A.RED_SUM, o[0], None), (A.RED_MAX, o[1], None), (A.RED_ARGMAX, o[2], o[4]), (A.RED_MEAN, o[3], None)]
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
        assert len({tuple(v.item() for v in o) for o in outs}) == 1   # every rank folded the same records in the same order
