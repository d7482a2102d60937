"""World-size-2 test of the multi-GPU host logic on CPU (gloo): slab partitioning, the 32-byte partial-record
layout, the single all-gather per step and the rank-order fold with lowest-GLOBAL-index ties.  The two device calls
(mxb_reduce_partial / mxb_reduce_finalize) are replaced by the CPU oracle writing / reading the same record bytes;
the exchange and the bookkeeping are the production code of matx_b200/dist.py."""
import os
import struct

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from matx_b200 import _abi as A
from matx_b200 import dist as mxd
from matx_b200 import ops as mx

N = 100_003


def test_slab_and_row_partitions_cover_exactly():
    for n in (1, 7, 1024, 1 << 20, (1 << 30)):
        for world in (1, 2, 3, 4, 8):
            parts = [mxd.slab(n, r, world) for r in range(world)]
            assert sum(c for _, c in parts) == n
            pos = 0
            for s, c in parts:
                assert s == pos or c == 0
                pos += c
                assert s % mxd.ALIGN == 0 or c == 0
            rows = [mxd.shard_rows(n, r, world) for r in range(world)]
            assert sum(c for _, c in rows) == n and max(c for _, c in rows) - min(c for _, c in rows) <= 1


class CpuSharded(mxd.ShardedFullReduce):
    """Same exchange, oracle in place of the two device calls."""

    def __init__(self, world, rank, oracle, x_local):
        self.oracle, self.x_local = oracle, x_local
        super().__init__(None, world, rank)

    def _alloc(self, nbytes):
        return torch.zeros(nbytes, dtype=torch.uint8)

    def _partial(self, op, operand, slab_offset, k):
        from tests.oracle_harness import np_tensor
        t = np_tensor(self.x_local)
        val = np.zeros((), np.float32)
        idx = np.zeros((), np.int64)
        if self.x_local.size == 0:
            raise AssertionError("empty slab in this test")
        self.oracle.reduce(mx.ReduceExpr(op, t, None), val, idx if op in (A.RED_ARGMAX, A.RED_ARGMIN) else None)
        rec = bytearray(32)
        struct.pack_into("<f", rec, 0, float(val))
        if op in (A.RED_ARGMAX, A.RED_ARGMIN):
            struct.pack_into("<q", rec, 8, int(idx) + slab_offset)
        if op in (A.RED_ANY, A.RED_ALL):
            struct.pack_into("<i", rec, 0, int(val != 0))
        self.records[k * 32:(k + 1) * 32] = torch.frombuffer(rec, dtype=torch.uint8)

    def _finalize(self, op, value_dtype, k, n_items, global_count, out, idx):
        raw = self.gathered.numpy().tobytes()
        recs = [raw[r * n_items * 32 + k * 32: r * n_items * 32 + (k + 1) * 32] for r in range(self.world)]
        vals = [struct.unpack_from("<f", r, 0)[0] for r in recs]
        if op == A.RED_SUM:
            acc = np.float32(0)
            for v in vals:
                acc = np.float32(acc + np.float32(v))
            out[...] = float(acc)
        elif op == A.RED_MEAN:
            acc = np.float32(0)
            for v in vals:
                acc = np.float32(acc + np.float32(v))
            out[...] = float(acc / np.float32(global_count))
        elif op == A.RED_MAX:
            out[...] = max(vals)
        elif op == A.RED_ARGMAX:
            ids = [struct.unpack_from("<q", r, 8)[0] for r in recs]
            best = max(vals)
            out[...] = best
            idx[...] = min(i for v, i in zip(vals, ids) if v == best)


def _worker(rank, world, port, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tests.oracle_harness import load_oracle
    oracle = load_oracle()
    x = np.random.default_rng(42).integers(0, 50, N).astype(np.float32)   # many ties at the maximum
    start, count = mxd.slab(N, rank, world, align=64)
    sh = CpuSharded(world, rank, oracle, x[start:start + count])
    o_sum, o_max, o_amax, o_idx = torch.zeros(()), torch.zeros(()), torch.zeros(()), torch.zeros((), dtype=torch.int64)
    items = [(A.RED_SUM, o_sum, None), (A.RED_MAX, o_max, None), (A.RED_ARGMAX, o_amax, o_idx)]
    sh.run(items, None, start, N, value_dtype=A.F32)
    ret[rank] = (float(o_sum), float(o_max), float(o_amax), int(o_idx))
    dist.destroy_process_group()


def test_sharded_full_reduce_world2_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    x = np.random.default_rng(42).integers(0, 50, N).astype(np.float32)
    assert ret[0] == ret[1]                                  # every rank folds the same records in the same order
    s, m, am, ix = ret[0]
    assert m == x.max() and am == x.max()
    assert ix == int(np.argmax(x))                           # lowest GLOBAL index among the tied maxima
    assert abs(s - x.astype(np.float64).sum()) <= 1e-5 * x.sum()
