"""Multi-GPU layer (no counterpart in the reference, which is single-device: SURVEY.md section 2.3 / 8e).

One process per GPU, `torch.distributed` for the plumbing.
  * batched reductions / elementwise: shard the outer (batch) dim with `shard_rows` — no communication at all;
  * full-tensor reductions: contiguous slab per rank (`slab`), each rank reduces its slab into 32-byte partial
    records with mxb_reduce_partial, ONE all-gather exchanges the records of every statement of the step, and
    every rank folds them in rank order with mxb_reduce_finalize (deterministic; the lowest GLOBAL index wins
    argmax / argmin ties because slab offsets are folded into the indices before the exchange).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

from . import _abi as A
from . import ops as mx

ALIGN = 1024  # slab boundaries stay aligned for 128/256-bit loads


def slab(n: int, rank: int, world: int, align: int = ALIGN) -> tuple[int, int]:
    """(start, count) of rank's contiguous slab of an n-element tensor; counts differ by at most `align`."""
    per = -(-n // world)
    per = -(-per // align) * align
    start = min(rank * per, n)
    return start, max(0, min(n, start + per) - start)


def shard_rows(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    """(start, count) of rank's block of the outermost batch dim (no communication needed afterwards)."""
    base, rem = divmod(n_rows, world)
    start = rank * base + min(rank, rem)
    return start, base + (1 if rank < rem else 0)


class ShardedFullReduce:
    """Full-tensor reductions over slab-sharded data: partial -> one all-gather -> finalize."""

    MAX_ITEMS = 8

    def __init__(self, ex, world: int, rank: int, group=None):
        import torch
        self.ex, self.world, self.rank, self.group = ex, world, rank, group
        self._torch = torch
        self.records = self._alloc(self.MAX_ITEMS * A.MXB_PARTIAL_BYTES)
        self.gathered = self._alloc(world * self.MAX_ITEMS * A.MXB_PARTIAL_BYTES)

    # -- overridable pieces (the CPU / gloo tests substitute the oracle for the two device calls) --
    def _alloc(self, nbytes: int):
        return self._torch.zeros(nbytes, dtype=self._torch.uint8, device="cuda")

    def _partial(self, op: int, operand, slab_offset: int, k: int) -> None:
        e = mx.lower_reduce(mx.ReduceExpr(op, operand, None))
        ptr = self.records.data_ptr() + k * A.MXB_PARTIAL_BYTES
        A.check(A.lib.mxb_reduce_partial(self.ex.handle, op, C.byref(e), slab_offset, C.c_void_p(ptr)))

    def _exchange(self, n_items: int) -> None:
        import torch.distributed as dist
        nb = n_items * A.MXB_PARTIAL_BYTES
        dist.all_gather_into_tensor(self.gathered[: self.world * nb], self.records[:nb], group=self.group)

    def _finalize(self, op: int, value_dtype: int, k: int, n_items: int, global_count: int, out, idx) -> None:
        stride = n_items * A.MXB_PARTIAL_BYTES
        ptr = self.gathered.data_ptr() + k * A.MXB_PARTIAL_BYTES
        o = mx._out_desc(mx.make_tensor(out))
        io = mx._out_desc(mx.make_tensor(idx)) if idx is not None else None
        A.check(A.lib.mxb_reduce_finalize(self.ex.handle, op, value_dtype, C.c_void_p(ptr), self.world, stride, global_count, 1,
                                          C.byref(o), C.byref(io) if io is not None else None))

    def run(self, items: Sequence[tuple], operand, slab_offset: int, global_count: int, value_dtype: Optional[int] = None) -> None:
        """items = [(reduce_op, out_tensor, idx_tensor_or_None), ...] all over the same sharded operand."""
        if len(items) > self.MAX_ITEMS:
            raise ValueError("at most %d statements per exchange" % self.MAX_ITEMS)
        if value_dtype is None:
            value_dtype = operand.dtype_hint if operand.dtype_hint not in (A.BF16, A.F16) else A.F32
        for k, (op, _, _) in enumerate(items):
            self._partial(op, operand, slab_offset, k)
        self._exchange(len(items))
        for k, (op, out, idx) in enumerate(items):
            self._finalize(op, value_dtype, k, len(items), global_count, out, idx)

    # -- prepared form: the statements are lowered once, a step is then 2 x len(items) C calls + one all-gather --
    def prepare(self, items: Sequence[tuple], operand, slab_offset: int, global_count: int, value_dtype: Optional[int] = None):
        if len(items) > self.MAX_ITEMS:
            raise ValueError("at most %d statements per exchange" % self.MAX_ITEMS)
        if value_dtype is None:
            value_dtype = operand.dtype_hint if operand.dtype_hint not in (A.BF16, A.F16) else A.F32
        n = len(items)
        stride = n * A.MXB_PARTIAL_BYTES
        plan = {"n": n, "partial": [], "final": [], "keep": []}
        for k, (op, out, idx) in enumerate(items):
            e = mx.lower_reduce(mx.ReduceExpr(op, operand, None))
            rec = C.c_void_p(self.records.data_ptr() + k * A.MXB_PARTIAL_BYTES)
            plan["partial"].append((op, e, C.c_int64(slab_offset), rec))
            o = mx._out_desc(mx.make_tensor(out))
            io = mx._out_desc(mx.make_tensor(idx)) if idx is not None else None
            g = C.c_void_p(self.gathered.data_ptr() + k * A.MXB_PARTIAL_BYTES)
            plan["final"].append((op, value_dtype, g, stride, C.c_int64(global_count), o, io))
            plan["keep"].append((out, idx, operand))
        return plan

    def run_prepared(self, plan) -> None:
        lib, h = A.lib, self.ex.handle
        for op, e, off, rec in plan["partial"]:
            A.check(lib.mxb_reduce_partial(h, op, C.byref(e), off, rec))
        self._exchange(plan["n"])
        for op, vdt, g, stride, gcount, o, io in plan["final"]:
            A.check(lib.mxb_reduce_finalize(h, op, vdt, g, self.world, stride, gcount, 1, C.byref(o), C.byref(io) if io is not None else None))


class PeerExchange:
    """Exchange buffers of every rank, mapped into this process through CUDA IPC, for the fused (NCCL-free) exchange:
    mxb_reduce_partial_push stores records into all ranks' buffers over NVLink, mxb_exchange_finalize waits and folds.
    `torch.distributed` is used once, to pass the IPC handles around."""

    FLAG_OFF_PAD = 256

    def __init__(self, ex, world: int, rank: int, group=None, _sim_buffers=None):
        import torch
        self.ex, self.world, self.rank = ex, world, rank
        rec_bytes = A.exchange_rec_bytes(world)
        self.rec_bytes = rec_bytes
        total = rec_bytes + 2 * self.FLAG_OFF_PAD
        self._keep = []
        if _sim_buffers is not None:              # single-process simulation of `world` ranks on one device (tests)
            bufs = _sim_buffers
            self.buf = bufs[rank]
            bases = [b.data_ptr() for b in bufs]
        else:
            import torch.distributed as dist
            own = C.c_void_p()
            hbuf = C.create_string_buffer(64)
            A.check(A.lib.mxb_exchange_alloc(ex.handle, total, C.byref(own), hbuf))
            self._own, self._opened = own, []
            handles = [None] * world
            dist.all_gather_object(handles, bytes(hbuf.raw), group=group)
            bases = []
            for r, hd in enumerate(handles):
                if r == rank:
                    bases.append(own.value)
                    continue
                pp = C.c_void_p()
                A.check(A.lib.mxb_exchange_open(ex.handle, hd, C.byref(pp)))
                self._opened.append(pp)
                bases.append(pp.value)
            dist.barrier(group=group)
            self.buf = None
            self._base = own.value
        self.peers = A.Peers()
        for r in range(world):
            self.peers.rec[r] = bases[r]
            self.peers.flag[r] = bases[r] + rec_bytes
        own_base = bases[rank]
        self.peers.epoch = own_base + rec_bytes + self.FLAG_OFF_PAD
        self.peers.world, self.peers.rank = world, rank

    def close(self) -> None:
        """Unmap the peers' buffers and free this rank's (call after a barrier: peers may still be reading)."""
        for pp in getattr(self, "_opened", []):
            A.lib.mxb_exchange_close(self.ex.handle, pp)
        self._opened = []
        if getattr(self, "_own", None) is not None:
            A.lib.mxb_exchange_free(self.ex.handle, self._own)
            self._own = None

    @staticmethod
    def buffer_bytes(world: int) -> int:
        return A.exchange_rec_bytes(world) + 2 * PeerExchange.FLAG_OFF_PAD

    def prepare(self, items: Sequence[tuple], operand, slab_offset: int, global_count: int, value_dtype: Optional[int] = None, ddof: int = 1):
        if len(items) > A.MXB_MAX_ITEMS:
            raise ValueError("at most %d statements per exchange" % A.MXB_MAX_ITEMS)
        if value_dtype is None:
            value_dtype = operand.dtype_hint if operand.dtype_hint not in (A.BF16, A.F16) else A.F32
        n = len(items)
        fold = (A.FoldItem * n)()
        pushes = []
        for k, (op, out, idx) in enumerate(items):
            e = mx.lower_reduce(mx.ReduceExpr(op, operand, None))
            pushes.append((op, e, C.c_int64(slab_offset), k))
            fold[k].reduce_op, fold[k].value_dtype = op, value_dtype
            fold[k].out = out.data_ptr()
            fold[k].idx_out = idx.data_ptr() if idx is not None else None
            fold[k].ddof = ddof
        return {"n": n, "push": pushes, "fold": fold, "count": C.c_int64(global_count), "keep": (items, operand)}

    def run_prepared(self, plan) -> None:
        lib, h, peers = A.lib, self.ex.handle, C.byref(self.peers)
        for op, e, off, k in plan["push"]:
            A.check(lib.mxb_reduce_partial_push(h, op, C.byref(e), off, peers, k, plan["n"]))
        A.check(lib.mxb_exchange_finalize(h, peers, plan["fold"], plan["n"], plan["count"]))

    def check(self) -> None:
        """Synchronise and raise if any exchange since the last check timed out on a peer (its outputs hold NaN / -1)."""
        A.check(A.lib.mxb_exchange_check(self.ex.handle, C.byref(self.peers)))
