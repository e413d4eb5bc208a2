"""Multi-GPU particle field: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch), particles
block-partitioned by index.

Direct path.  Each rank owns the targets [lo, hi) and their full state; sources are a read-only broadcast.  Per U/J
evaluation a rank packs its particles into source tiles (the engine's wire format, 2570 doubles per 256 sources), the tiles
are all-gathered, and the pair kernel runs once per peer's tile set.  The gather is issued asynchronously and the rank's OWN
tiles are processed first, so the NVLink transfer (56-80 MB at N = 1M, ~0.1 ms at 770 GB/s) is hidden behind the first
of `world` kernel launches; nothing else in the step communicates (update / SFS-coefficient / relaxation kernels are
local to the owner shard).

UJ_fmm.  A local essential tree (`fmm="let"`, default): the Morton curve is cut into `world` ranges at the unit boundaries
of the global tree's top, particles travel to the owner of their range (all-to-all), every rank builds ITS part of the one
global octree and runs the upward pass on it, tree skeletons / multipoles / source records are exchanged, every rank
evaluates its own targets against all trees with the single-GPU kernels, and the results travel home with the inverse
all-to-all (per-rank phases: flowunsteady_b200/csrc/fmm_let.cuh).  Sort, tree build, upward pass, traversal, M2L and near
field are all distributed; results equal the one-GPU UJ_fmm to round-off.  `fmm="replicated"` keeps round 1's scheme
(every rank builds the whole tree, leaves are split, one all-reduce) as a cross-check.
The reference has no distributed mode at all (README.md:145 of the reference).

The orchestration is written against a small backend protocol (pack / from_records / stage / let_*) and a collectives
object, so the world_size-2 gloo tests on CPU can drive the direct path with a numpy stand-in and the LET path can be
tested with several engines on ONE GPU (tests/loopback.py); the product backend is `Engine` (CUDA) + `TorchCollectives`.
"""
from __future__ import annotations

import contextlib
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import engine as _E

RK3 = ((0.0, 1.0 / 3.0), (-5.0 / 9.0, 15.0 / 16.0), (-153.0 / 128.0, 8.0 / 15.0))
LET_LEVEL = 5          # level of the histogram that cuts the Morton curve: 8^5 = 32,768 bins


class _DevArray:
    """Raw CUDA pointer -> torch (via __cuda_array_interface__); the engine keeps ownership."""

    def __init__(self, ptr: int, shape, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False), "version": 3}


def partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous index ranges, sizes differing by at most one particle."""
    base, rem = divmod(int(n), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class TorchCollectives:
    """The collectives the sharded field needs, on torch.distributed (NCCL on GPUs, gloo in the CPU tests)."""

    def __init__(self, group=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def all_reduce_(self, t: torch.Tensor, op: str = "sum"):
        if self.world > 1:
            dist.all_reduce(t, op={"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}[op], group=self.group)
        return t

    def all_gather(self, t: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        """(world, *t.shape): every rank's `t` (equal shapes); `out`: optional flat destination of world * t.numel()."""
        if self.world == 1:
            return t.unsqueeze(0)
        if out is None:
            out = torch.empty(self.world * t.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out.view(-1), t.contiguous().view(-1), group=self.group)
        return out.view((self.world,) + tuple(t.shape))

    def all_gather_into_async(self, out: torch.Tensor, t: torch.Tensor):
        if self.world == 1:
            return None
        return dist.all_gather_into_tensor(out, t, group=self.group, async_op=True)

    def all_gather_ints(self, vals: Sequence[int], device) -> List[List[int]]:
        t = torch.tensor([int(v) for v in vals], dtype=torch.int64, device=device)
        return [[int(x) for x in row] for row in self.all_gather(t).tolist()]

    def all_to_all_rows(self, send: torch.Tensor, send_counts: Sequence[int], recv_counts: Sequence[int],
                        out: torch.Tensor | None = None) -> torch.Tensor:
        """Rows [sum(send_counts[:k]), +send_counts[k]) of `send` go to rank k; returns the rows received, by source rank
        (`out`: optional destination with at least sum(recv_counts) rows)."""
        n = int(sum(recv_counts))
        if self.world == 1:
            return send[:n]
        recv = out if out is not None else torch.empty((max(n, 1),) + tuple(send.shape[1:]), dtype=send.dtype, device=send.device)
        dist.all_to_all_single(recv[:n], send[:int(sum(send_counts))], [int(c) for c in recv_counts], [int(c) for c in send_counts],
                               group=self.group)
        return recv[:n]


class ShardedField:
    """Drives pfield.UJ / pfield.SFS / vpm.nextstep over a particle field sharded across the ranks of `group`.

    backend: object with the Engine methods used below (np, tiles_for, tile_doubles, pack_uj_records,
             pack_estr_records, uj_from_records, estr_from_records, stage, get_schemes, set_time, get_time, stream, let_*).
    device:  torch device the exchange buffers live on ("cuda:N" for Engine).
    coll:    collectives object (default: TorchCollectives(group)).
    fmm:     "let" (local essential tree, every rank receives all skeletons, multipoles and records), "let_halo" (same tree
             and same results, but only the skeletons are all-gathered: multipoles and source records travel on demand,
             owner -> the ranks whose traversal listed them; per-rank memory ~ own share + halo) or "replicated" (round-1
             scheme) for vpm_UJ = UJ_fmm.
    """

    def __init__(self, backend, max_local: int, device, group=None, coll=None, fmm: str = "let", let_level: int = LET_LEVEL):
        self.b = backend
        self.group = group
        self.coll = coll if coll is not None else TorchCollectives(group)
        self.world, self.rank = self.coll.world, self.coll.rank
        self.device = torch.device(device)
        if fmm not in ("let", "let_halo", "replicated"):
            raise ValueError(f"fmm must be 'let', 'let_halo' or 'replicated', not {fmm!r}")
        self.fmm_mode = fmm if hasattr(backend, "let_bounds") else "replicated"
        self.let_level = int(let_level)
        self.td = int(backend.tile_doubles())
        # every rank's slot in the gathered buffer has the same capacity (all_gather needs equal counts)
        cap = torch.tensor([int(backend.tiles_for(max_local))], dtype=torch.int64, device=self._ctl_device())
        self.coll.all_reduce_(cap, "max")
        self.slot_tiles = int(cap.item())
        self.local = torch.zeros(self.slot_tiles * self.td, dtype=torch.float64, device=self.device)
        self.gathered = (torch.zeros(self.world * self.slot_tiles * self.td, dtype=torch.float64, device=self.device)
                         if self.world > 1 else self.local)
        self._ntiles = [0] * self.world
        self._fmm_hint = 0          # 1: the next UJ_fmm evaluation's far field may be reused; 2: reuse it (DynamicSFS)
        self._let = None            # partition / exchange state of the last LET evaluation
        self.let_timing = None      # set to {} to collect wall-clock milliseconds per LET phase (synchronising; diagnostics)
        self._bufs = {}             # grow-only exchange buffers of the LET path (no allocator traffic inside a step)
        self.let_balance = True     # cut the Morton curve by counted WORK (per-bin interaction counts of the previous evaluation)
        self._let_work_ready = False
        self.refresh_counts()

    def _ctl_device(self):
        return self.device if self.device.type == "cuda" else torch.device("cpu")

    def refresh_counts(self):
        """All ranks learn every rank's tile count (call after the local particle count changes)."""
        mine = int(self.b.tiles_for(self.b.np))
        if mine > self.slot_tiles:
            raise RuntimeError("local shard outgrew the tile slot; recreate the ShardedField with a larger max_local")
        t = torch.zeros(self.world, dtype=torch.int64, device=self._ctl_device())
        t[self.rank] = mine
        self.coll.all_reduce_(t, "sum")
        self._ntiles = [int(v) for v in t.tolist()]
        self._let = None

    # ---- stream plumbing: collectives are ordered against the backend's CUDA stream -----------------------------
    def _stream_ctx(self):
        if self.device.type == "cuda":
            return torch.cuda.stream(torch.cuda.ExternalStream(self.b.stream, device=self.device))
        return contextlib.nullcontext()

    def _slot_ptr(self, r: int) -> int:
        return self.gathered.data_ptr() + r * self.slot_tiles * self.td * 8

    def _pairwise(self, pack, apply_first, apply_rest):
        """pack -> async all-gather -> own tiles -> wait -> every peer's tiles."""
        with self._stream_ctx():
            pack(self.local.data_ptr())
            work = self.coll.all_gather_into_async(self.gathered, self.local)
            apply_first(self.local.data_ptr(), self._ntiles[self.rank])
            if work is not None:
                work.wait()
                for r in range(self.world):
                    if r != self.rank and self._ntiles[r] > 0:
                        apply_rest(self._slot_ptr(r), self._ntiles[r])

    def _state_view(self):
        """(43, ld) torch view of the backend's SoA state (no copy).  The CUDA engine hands out its device pointer; the
        numpy stand-in of the gloo tests provides `state_tensor()` itself."""
        if hasattr(self.b, "state_tensor"):
            return self.b.state_tensor()
        ptr, ld = self.b.device_field(0)
        return torch.as_tensor(_DevArray(ptr, (43, ld)), device=self.device)

    # ---- UJ_fmm over the sharded field: local essential tree ---------------------------------------------------------
    def _dev_view(self, ptr: int, n: int, typestr: str = "<f8", dtype=torch.float64):
        if n <= 0 or not ptr:
            return torch.empty(0, dtype=dtype, device=self.device)
        if self.device.type != "cuda":                # the numpy stand-in of the gloo tests hands out host pointers
            return self.b.host_view(ptr, n, dtype)
        return torch.as_tensor(_DevArray(ptr, (n,), typestr), device=self.device)

    def _buf(self, name: str, n: int, dtype=torch.float64) -> torch.Tensor:
        """Flat grow-only device buffer: sizes drift by a few particles / cells from one evaluation to the next, and a fresh
        torch allocation per evaluation (1.5 GB in all at 5M particles) would churn the caching allocator."""
        t = self._bufs.get(name)
        if t is None or t.numel() < n or t.dtype != dtype:
            t = torch.empty(int(n) + int(n) // 8 + 1024, dtype=dtype, device=self.device)
            self._bufs[name] = t
        return t[:int(n)]

    def _uj_fmm_let(self, reset: bool, reset_sfs: bool, sfs: bool):
        """One UJ_fmm evaluation with a local essential tree; phases and what is exchanged between them: fmm_let.cuh."""
        if sfs and not reset:
            raise NotImplementedError("sharded UJ_fmm with sfs=True needs reset=True (what every SFS scheme calls)")
        b, c, dev, G, r = self.b, self.coll, self.device, self.world, self.rank
        hint, self._fmm_hint = self._fmm_hint, 0
        sch = b.get_schemes()
        L = self._let
        import time as _time
        _t = [_time.perf_counter()]

        def lap(name):
            """phase timer (diagnostics only: it drains the stream)"""
            if self.let_timing is None:
                return
            b.synchronize()
            if dev.type == "cuda":
                torch.cuda.synchronize(dev)
            now = _time.perf_counter()
            self.let_timing[name] = self.let_timing.get(name, 0.0) + (now - _t[0]) * 1e3
            _t[0] = now

        reuse = bool(hint == 2 and L is not None and L.get("far_valid") and not sch.fmm_nonzero_sigma and L["n_home"] == int(b.np))
        with self._stream_ctx():
            n_home = int(b.np)
            if not reuse:
                lohi = torch.tensor(b.let_bounds(), dtype=torch.float64, device=dev)
                lohi[:3] = -lohi[:3]
                c.all_reduce_(lohi, "max")                           # one collective: max of (-min, max)
                lohi[:3] = -lohi[:3]
                lohi_g = lohi.cpu().numpy()
                L = self._let = {"n_home": n_home, "far_valid": False}
                if not (lohi_g[0] <= lohi_g[3]):                     # no particle anywhere
                    return
                hist_ptr, binmax_ptr = b.let_keys(lohi_g, self.let_level)
                bins = 8 ** self.let_level
                c.all_reduce_(self._dev_view(hist_ptr, bins, "<i4", torch.int32), "sum")
                if binmax_ptr:
                    c.all_reduce_(self._dev_view(binmax_ptr, bins), "max")
                use_work = bool(self.let_balance and self._let_work_ready and G > 1)
                L["send"] = b.let_partition(G, r, use_work)
                counts = c.all_gather_ints(L["send"], dev)           # counts[q][k]: particles rank q sends to rank k
                L["recv"] = [counts[q][r] for q in range(G)]
                L["n_all"] = sum(sum(row) for row in counts)
                lap("1 bounds, keys, histogram, partition")
            elif "send" not in L:
                return
            n_own = sum(L["recv"])
            send = self._buf("send", max(n_home, 1) * 7).view(-1, 7)
            b.let_pack(send.data_ptr())
            rows = c.all_to_all_rows(send[:n_home], L["send"], L["recv"], out=self._buf("rows", max(n_own, 1) * 7).view(-1, 7))
            L["rows"] = rows                                         # the engine reads it until the evaluation ends
            lap("2 pack + all-to-all of particle rows")
            halo = self.fmm_mode == "let_halo"
            info = b.let_build(rows.data_ptr(), n_own, n_own if halo else L["n_all"], reuse)
            cells_ptr, M_ptr, rec_ptr = b.let_ptrs()
            lap("3 owner sort, tree, upward pass")
            if halo:
                return self._let_halo_evaluate(L, info, cells_ptr, n_own, n_home, reuse, reset, reset_sfs, sfs, hint, sch, lap)
            if not reuse:
                ncells_own, _, nm3, _ = info
                sizes = c.all_gather_ints([n_own, ncells_own], dev)
                L["np"], L["nc"] = [s[0] for s in sizes], [s[1] for s in sizes]
                L["slot_c"], L["slot_n"] = max(max(L["nc"]), 1), max(max(L["np"]), 1)
                cb = b.let_cell_bytes()
                # slots are padded (the padding is never read: attach copies ncells[q] / np[q] entries of every block)
                sc = self._buf("sc", L["slot_c"] * cb, torch.uint8)
                sc[:ncells_own * cb] = self._dev_view(cells_ptr, ncells_own * cb, "|u1", torch.uint8)
                sm = self._buf("sm", L["slot_c"] * nm3)
                sm[:ncells_own * nm3] = self._dev_view(M_ptr, ncells_own * nm3)
                cells_all = c.all_gather(sc, out=self._buf("cells_all", G * L["slot_c"] * cb, torch.uint8))
                M_all = c.all_gather(sm, out=self._buf("M_all", G * L["slot_c"] * nm3))
                b.let_attach_tree(cells_all.data_ptr(), M_all.data_ptr(), L["slot_c"], L["nc"], L["np"])
                lap("4 all-gather skeletons + multipoles, attach")
            srec = self._buf("srec", L["slot_n"] * 10)

            def exchange_records(overlap=None):
                """all-gather of the owners' source records (UJ or E_str flavour) behind the own ones; `overlap`: device work
                that needs the skeletons only, enqueued while the records are in flight"""
                srec[:n_own * 10] = self._dev_view(rec_ptr, n_own * 10)
                rec_all = self._buf("rec_all", G * L["slot_n"] * 10)
                work = None
                if G == 1:
                    rec_all = srec
                elif overlap is not None:
                    work = c.all_gather_into_async(rec_all, srec)     # (a blocking gather with the in-process collectives)
                else:
                    c.all_gather(srec, out=rec_all)
                if overlap is not None:
                    overlap()
                if work is not None:
                    work.wait()
                b.let_attach_records(rec_all.data_ptr(), L["slot_n"], L["np"])
                return rec_all

            out = self._buf("out", max(n_own, 1) * 12).view(-1, 12)
            keep = exchange_records(overlap=lambda: b.let_evaluate(out.data_ptr(), reuse, 1))   # traversal, M2L, L2L meanwhile
            lap("5 all-gather source records || lists, M2L, L2L")
            b.let_evaluate(out.data_ptr(), reuse, 2)
            if not reuse and self.let_balance and G > 1:
                # the interaction work this evaluation counted per Morton bin, summed over the ranks: the next cut equalises it
                c.all_reduce_(self._dev_view(b.let_work(), 8 ** self.let_level, "<i8", torch.int64), "sum")
                self._let_work_ready = True
            lap("6 L2P + near field")
            res = c.all_to_all_rows(out[:n_own], L["recv"], L["send"], out=self._buf("res", max(n_home, 1) * 12).view(-1, 12))
            b.let_finish(res.data_ptr(), 0, reset)
            lap("7 inverse all-to-all of U, J + scatter")
            if reset_sfs:
                b.reset_particles_sfs()
            if sfs:
                b.let_estr_records()
                keep = exchange_records()
                outE = self._buf("outE", max(n_own, 1) * 3).view(-1, 3)
                b.let_estr_evaluate(outE.data_ptr())
                resE = c.all_to_all_rows(outE[:n_own], L["recv"], L["send"], out=self._buf("resE", max(n_home, 1) * 3).view(-1, 3))
                b.let_finish(resE.data_ptr(), 1, False)
                lap("8 E_str: records, all-gather, near field, return")
            L["far_valid"] = bool(hint == 1 and not sch.fmm_nonzero_sigma)
            del keep

    def _let_halo_evaluate(self, L, info, cells_ptr, n_own, n_home, reuse, reset, reset_sfs, sfs, hint, sch, lap):
        """Phases 4-8 of the LET evaluation with a demand-driven halo (vpmb200.h: vpmb200_let_attach_skeleton ...): the
        skeletons are all-gathered, the traversal runs, and every rank then REQUESTS the multipoles of the remote cells on
        its M2L list and the records of the remote leaves on its P2P list from their owners (two all-to-alls of ids, two of
        payload).  Called inside the stream context of _uj_fmm_let."""
        b, c, dev, G, r = self.b, self.coll, self.device, self.world, self.rank
        ncells_own, nleaves, nm3, _ = info
        out = self._buf("out", max(n_own, 1) * 12).view(-1, 12)
        if not reuse:
            sizes = c.all_gather_ints([n_own, ncells_own, nleaves if n_own > 0 else 0], dev)
            L["np"], L["nc"], L["nl"] = ([s[k] for s in sizes] for k in range(3))
            L["slot_c"] = max(max(L["nc"]), 1)
            cb = b.let_cell_bytes()
            sc = self._buf("sc", L["slot_c"] * cb, torch.uint8)
            sc[:ncells_own * cb] = self._dev_view(cells_ptr, ncells_own * cb, "|u1", torch.uint8)
            cells_all = c.all_gather(sc, out=self._buf("cells_all", G * L["slot_c"] * cb, torch.uint8))
            b.let_attach_skeleton(cells_all.data_ptr(), L["slot_c"], L["nc"], L["np"], L["nl"])
            lap("4 all-gather skeletons, attach")
            b.let_evaluate(out.data_ptr(), False, 5)                  # interaction lists
            want, pc, pl = b.let_halo_plan(G)                         # want[3 q + (0, 1, 2)]: cells, leaves, records from rank q
            allw = c.all_gather_ints(want, dev)                       # allw[q][3 k + j]: what rank q wants from rank k
            give = [allw[q][3 * r:3 * r + 3] for q in range(G)]       # what rank q wants from this rank
            wc, wl = [want[3 * q] for q in range(G)], [want[3 * q + 1] for q in range(G)]
            ids_c = self._dev_view(pc, max(sum(wc), 1), "<i4", torch.int32)[:sum(wc)].view(-1, 1)
            ids_l = self._dev_view(pl, 2 * max(sum(wl), 1), "<i4", torch.int32)[:2 * sum(wl)].view(-1, 2)
            gc, gl = [g[0] for g in give], [g[1] for g in give]
            L["ask_c"] = c.all_to_all_rows(ids_c, wc, gc, out=self._buf("ask_c", max(sum(gc), 1), torch.int32).view(-1, 1))
            L["ask_l"] = c.all_to_all_rows(ids_l, wl, gl, out=self._buf("ask_l", 2 * max(sum(gl), 1), torch.int32).view(-1, 2))
            L["want"], L["give"] = want, give
            lap("5 lists, halo plan, all-to-all of the requests")
        want, give = L["want"], L["give"]
        gc, gl, gr = ([g[k] for g in give] for k in range(3))
        wc, wr = [want[3 * q] for q in range(G)], [want[3 * q + 2] for q in range(G)]

        def serve(with_multipoles: bool):
            """owner side: gather what the others listed; then the payload all-to-all(s); then point the kernels at the halo"""
            M_out = self._buf("halo_M_out", max(sum(gc), 1) * nm3).view(-1, nm3) if with_multipoles else None
            rec_out = self._buf("halo_rec_out", max(sum(gr), 1) * 10).view(-1, 10)
            b.let_halo_serve(L["ask_c"].data_ptr(), sum(gc) if with_multipoles else 0, L["ask_l"].data_ptr(), sum(gl),
                             M_out.data_ptr() if with_multipoles else 0, rec_out.data_ptr())
            if with_multipoles:
                L["M2"] = c.all_to_all_rows(M_out[:sum(gc)], gc, wc, out=self._buf("halo_M", max(sum(wc), 1) * nm3).view(-1, nm3))
            L["rec2"] = c.all_to_all_rows(rec_out[:sum(gr)], gr, wr, out=self._buf("halo_rec", max(sum(wr), 1) * 10).view(-1, 10))
            b.let_halo_set(L["M2"].data_ptr(), L["rec2"].data_ptr())
            L["halo_bytes"] = 8 * (sum(wc) * nm3 + sum(wr) * 10)

        serve(not reuse)
        lap("6 halo: serve, all-to-all of multipoles + records")
        b.let_evaluate(out.data_ptr(), reuse, 6)                      # M2L + L2L
        b.let_evaluate(out.data_ptr(), reuse, 2)                      # L2P + near field + output rows
        if not reuse and self.let_balance and G > 1:
            c.all_reduce_(self._dev_view(b.let_work(), 8 ** self.let_level, "<i8", torch.int64), "sum")
            self._let_work_ready = True
        lap("7 M2L, L2L, L2P + near field")
        res = c.all_to_all_rows(out[:n_own], L["recv"], L["send"], out=self._buf("res", max(n_home, 1) * 12).view(-1, 12))
        b.let_finish(res.data_ptr(), 0, reset)
        lap("8 inverse all-to-all of U, J + scatter")
        if reset_sfs:
            b.reset_particles_sfs()
        if sfs:
            b.let_estr_records()
            serve(False)
            outE = self._buf("outE", max(n_own, 1) * 3).view(-1, 3)
            b.let_estr_evaluate(outE.data_ptr())
            resE = c.all_to_all_rows(outE[:n_own], L["recv"], L["send"], out=self._buf("resE", max(n_home, 1) * 3).view(-1, 3))
            b.let_finish(resE.data_ptr(), 1, False)
            lap("9 E_str: records, halo, near field, return")
        L["far_valid"] = bool(hint == 1 and not sch.fmm_nonzero_sigma)

    # ---- UJ_fmm over the sharded field: replicated tree, leaves split over the ranks (round-1 scheme) ----------------
    def _uj_fmm(self, reset: bool, reset_sfs: bool, sfs: bool):
        """All ranks gather (X, Gamma, sigma) of every particle (56 B each), build the SAME tree, evaluate 1/world of the
        leaves (near field + L2P are > 85 % of the evaluation), and combine with one all-reduce of the U, J rows (the
        rows of particles outside a rank's share are zeros).  E_str: a second near-field pass + all-reduce."""
        if sfs and not reset:
            raise NotImplementedError("sharded UJ_fmm with sfs=True needs reset=True (what every SFS scheme calls)")
        b, dev = self.b, self.device
        n_loc = int(b.np)
        cnt = torch.zeros(self.world, dtype=torch.int64, device=dev)
        cnt[self.rank] = n_loc
        self.coll.all_reduce_(cnt, "sum")
        counts = [int(v) for v in cnt.tolist()]
        ntot, slot = sum(counts), max(max(counts), 1)
        off = sum(counts[:self.rank])
        with self._stream_ctx():
            state = self._state_view()
            ldg = (ntot + 31) // 32 * 32
            if getattr(self, "_G", None) is None or self._G.shape[1] < ldg:
                self._G = torch.zeros((24, ldg), dtype=torch.float64, device=dev)
            G = self._G
            send = torch.zeros((7, slot), dtype=torch.float64, device=dev)
            send[:, :n_loc] = state[0:7, :n_loc]
            if self.world > 1:
                recv = self.coll.all_gather(send)
                o = 0
                for r, c in enumerate(counts):
                    G[0:7, o:o + c] = recv[r, :, :c]
                    o += c
            else:
                G[0:7, :n_loc] = send[:, :n_loc]
            b.fmm_global(G.data_ptr(), G.shape[1], ntot, self.rank, self.world, 0)
            self.coll.all_reduce_(G[9:24], "sum")
            mine = slice(off, off + n_loc)
            if reset:
                state[9:12, :n_loc] = G[9:12, mine]
                state[15:24, :n_loc] = G[15:24, mine]
                state[24:27, :n_loc] = 0.0
            else:
                state[9:12, :n_loc] += G[9:12, mine]
                state[15:24, :n_loc] += G[15:24, mine]
            if reset_sfs:
                state[39:42, :n_loc] = 0.0
            if sfs:
                b.fmm_global(G.data_ptr(), G.shape[1], ntot, self.rank, self.world, 1)
                self.coll.all_reduce_(G[12:15], "sum")
                state[39:42, :n_loc] += G[12:15, mine]

    # ---- pfield.UJ(pfield; reset, reset_sfs, sfs) -------------------------------------------------------------------
    def uj(self, reset: bool = True, reset_sfs: bool = False, sfs: bool = False):
        b = self.b
        if b.get_schemes().uj == _E.UJ_IDS["fmm"]:
            if self.fmm_mode in ("let", "let_halo"):
                return self._uj_fmm_let(reset, reset_sfs, sfs)
            return self._uj_fmm(reset, reset_sfs, sfs)
        if reset:
            b.reset_particles()        # U, J, PSE <- 0 (the pair kernel then accumulates chunk by chunk)
        if reset_sfs:
            b.reset_particles_sfs()
        self._pairwise(b.pack_uj_records,
                       lambda p, n: b.uj_from_records(p, n, True),
                       lambda p, n: b.uj_from_records(p, n, True))
        if sfs:
            self._pairwise(b.pack_estr_records, b.estr_from_records, b.estr_from_records)

    # ---- pfield.SFS(pfield; a, b)  (mirrors engine.cu do_sfs) -------------------------------------------------------
    def sfs(self, a: float = 1.0, b_: float = 1.0):
        s = self.b.get_schemes()
        first = a in (0.0, 1.0)
        st = self.b.stage
        if s.sfs == _E.SFS_IDS["none"]:
            self.uj(True, False, False)
        elif s.sfs == _E.SFS_IDS["constant"]:
            self.uj(True, True, True)
            if first:
                st(_E.STAGE_CONSTANT_COEFF)
                st(_E.STAGE_CLIP_CONTROL)
        else:
            if not first:
                self.uj(True, True, True)
                return
            st(_E.STAGE_SCALE_SIGMA_TEST)
            self._fmm_hint = 1      # UJ_fmm: keep the tree, the lists and the far field ...
            self.uj(True, True, True)
            st(_E.STAGE_STORE_TEST)
            st(_E.STAGE_SCALE_SIGMA_DOMAIN)
            self._fmm_hint = 2      # ... only sigma changed in between (engine.cu do_sfs does the same on one GPU)
            self.uj(True, True, True)
            self._fmm_hint = 0
            if self._let is not None:
                self._let["far_valid"] = False
            st(_E.STAGE_DYNAMIC_COEFF)
            st(_E.STAGE_CLIP_CONTROL)

    # ---- vpm.nextstep(pfield, dt; relax)  (mirrors engine.cu vpmb200_nextstep) ---------------------------------------
    def nextstep(self, dt: float, Uinf: Sequence[float] = (0.0, 0.0, 0.0), relax: bool = True):
        s = self.b.get_schemes()
        st = self.b.stage
        if s.viscous == _E.VISCOUS_IDS["corespreading"] and getattr(s, "cs_sgm0", 0.0) > 0:
            raise NotImplementedError("CoreSpreading's RBF re-fit is single-GPU; run it with cs_sgm0 = 0 on a sharded field")
        if sum(self._ntiles) > 0:
            if s.integration == _E.INTEGRATION_IDS["euler"]:
                self.sfs(1.0, 1.0)
                st(_E.STAGE_UPDATE_EULER_RELAX if relax else _E.STAGE_UPDATE, 0.0, 1.0, dt, Uinf)
            else:
                st(_E.STAGE_ZERO_M)
                for a, b_ in RK3:
                    self.sfs(a, b_)
                    st(_E.STAGE_UPDATE, a, b_, dt, Uinf)
                if relax and s.relaxation != _E.RELAX_IDS["none"]:
                    self.uj(True, False, False)
                    st(_E.STAGE_RELAX)
        t, nt = self.b.get_time()
        self.b.set_time(t + dt, nt + 1)
