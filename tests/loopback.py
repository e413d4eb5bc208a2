"""In-process collectives: `world` ranks run as Python THREADS of one process, each driving its own engine on the SAME GPU.

Test infrastructure for the multi-GPU UJ_fmm (local essential tree, flowunsteady_b200/dist.py): the per-rank phases and the
orchestration are exactly the product's; only the collectives are replaced (shared slots + a barrier instead of NCCL), so the
whole exchange logic — all-to-all of particle rows, all-gather of skeletons / multipoles / records, the inverse all-to-all —
is exercised on the one-GPU box the round-end test run uses.  The NCCL flavour is covered by tests/test_gpu_dist.py.
"""
from __future__ import annotations

import threading
from typing import List, Sequence

import torch


class LoopbackWorld:
    def __init__(self, world: int):
        self.world = world
        self.barrier = threading.Barrier(world)
        self.slots = [None] * world


class LoopbackCollectives:
    """Same interface as flowunsteady_b200.dist.TorchCollectives."""

    def __init__(self, W: LoopbackWorld, rank: int):
        self.W, self.rank, self.world = W, rank, W.world

    def _exchange(self, obj):
        """Every rank deposits `obj`; returns the list of all ranks' objects (valid until the NEXT collective)."""
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()     # the deposit must be complete before another thread reads it
        self.W.barrier.wait()                             # nobody is still reading the previous round's slots
        self.W.slots[self.rank] = obj
        self.W.barrier.wait()
        return list(self.W.slots)

    def _done_reading(self):
        if torch.cuda.is_available():
            torch.cuda.current_stream().synchronize()
        self.W.barrier.wait()

    def all_reduce_(self, t: torch.Tensor, op: str = "sum"):
        allv = self._exchange(t)
        st = torch.stack([v.to(t.device) for v in allv])
        res = {"sum": st.sum(0), "max": st.max(0).values, "min": st.min(0).values}[op].to(t.dtype)
        self._done_reading()
        t.copy_(res)
        return t

    def all_gather(self, t: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        allv = self._exchange(t)
        res = torch.stack([v.to(t.device) for v in allv])
        if out is not None:
            out.view(-1)[:res.numel()].copy_(res.view(-1))
            res = out.view(-1)[:res.numel()].view(res.shape)
        self._done_reading()
        return res

    def all_gather_into_async(self, out: torch.Tensor, t: torch.Tensor):
        out.view(self.world, -1).copy_(self.all_gather(t).view(self.world, -1))
        return None

    def all_gather_ints(self, vals: Sequence[int], device) -> List[List[int]]:
        allv = self._exchange([int(v) for v in vals])
        out = [list(v) for v in allv]
        self.W.barrier.wait()
        return out

    def all_to_all_rows(self, send: torch.Tensor, send_counts: Sequence[int], recv_counts: Sequence[int], out=None) -> torch.Tensor:
        offs = [0]
        for c in send_counts:
            offs.append(offs[-1] + int(c))
        allv = self._exchange((send, offs))
        parts = []
        for q, (s, o) in enumerate(allv):
            assert o[self.rank + 1] - o[self.rank] == int(recv_counts[q])
            parts.append(s[o[self.rank]:o[self.rank + 1]].to(send.device))
        res = torch.cat(parts) if parts else send[:0]
        if out is not None:
            out[:res.shape[0]].copy_(res)
            res = out[:res.shape[0]]
        self._done_reading()
        return res


def run_ranks(world: int, fn):
    """fn(rank, collectives) on `world` threads; re-raises the first exception; returns the per-rank results."""
    W = LoopbackWorld(world)
    results, errors = [None] * world, []

    def body(r):
        try:
            results[r] = fn(r, LoopbackCollectives(W, r))
        except BaseException as exc:   # noqa: BLE001 - propagate to the main thread
            errors.append(exc)
            W.barrier.abort()

    threads = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        real = [e for e in errors if not isinstance(e, threading.BrokenBarrierError)]
        raise (real or errors)[0]
    return results
