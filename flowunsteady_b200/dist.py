"""Multi-GPU direct path: one process per GPU (torch.distributed, NCCL over NVLink/NVSwitch), particles
block-partitioned by index.

Each rank owns the targets [lo, hi) and their full state; sources are a read-only broadcast.  Per U/J evaluation a
rank packs its particles into source tiles (the engine's wire format, 2570 doubles per 256 sources), the tiles are
all-gathered, and the pair kernel runs once per peer's tile set.  The gather is issued asynchronously and the rank's OWN
tiles are processed first, so the NVLink transfer (56-80 MB at N = 1M, ~0.1 ms at 770 GB/s) is hidden behind the first
of `world` kernel launches; nothing else in the step communicates (update / SFS-coefficient / relaxation kernels are
local to the owner shard).  The reference has no distributed mode at all (README.md:145 of the reference).

The orchestration is written against a small backend protocol (pack / from_records / stage) so the world_size-2 gloo
tests on CPU can drive it with a numpy stand-in; the product backend is `Engine` (CUDA).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist

from . import engine as _E

RK3 = ((0.0, 1.0 / 3.0), (-5.0 / 9.0, 15.0 / 16.0), (-153.0 / 128.0, 8.0 / 15.0))


class _DevArray:
    """Raw CUDA pointer -> torch (via __cuda_array_interface__); the engine keeps ownership."""

    def __init__(self, ptr: int, shape):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": "<f8", "data": (int(ptr), False), "version": 3}


def partition(n: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous index ranges, sizes differing by at most one particle."""
    base, rem = divmod(int(n), int(world))
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < rem else 0)
        out.append((lo, hi))
        lo = hi
    return out


class ShardedField:
    """Drives pfield.UJ / pfield.SFS / vpm.nextstep over a particle field sharded across the ranks of `group`.

    backend: object with the Engine methods used below (np, tiles_for, tile_doubles, pack_uj_records,
             pack_estr_records, uj_from_records, estr_from_records, stage, get_schemes, set_time, get_time, stream).
    device:  torch device the tile buffers live on ("cuda:N" for Engine).
    """

    def __init__(self, backend, max_local: int, device, group=None):
        self.b = backend
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.device = torch.device(device)
        self.td = int(backend.tile_doubles())
        # every rank's slot in the gathered buffer has the same capacity (all_gather needs equal counts)
        cap = torch.tensor([int(backend.tiles_for(max_local))], dtype=torch.int64, device=self._ctl_device())
        if self.world > 1:
            dist.all_reduce(cap, op=dist.ReduceOp.MAX, group=group)
        self.slot_tiles = int(cap.item())
        self.local = torch.zeros(self.slot_tiles * self.td, dtype=torch.float64, device=self.device)
        self.gathered = (torch.zeros(self.world * self.slot_tiles * self.td, dtype=torch.float64, device=self.device)
                         if self.world > 1 else self.local)
        self._ntiles = [0] * self.world
        self.refresh_counts()

    def _ctl_device(self):
        return self.device if self.device.type == "cuda" else torch.device("cpu")

    def refresh_counts(self):
        """All ranks learn every rank's tile count (call after the local particle count changes)."""
        mine = int(self.b.tiles_for(self.b.np))
        if mine > self.slot_tiles:
            raise RuntimeError("local shard outgrew the tile slot; recreate the ShardedField with a larger max_local")
        if self.world > 1:
            t = torch.zeros(self.world, dtype=torch.int64, device=self._ctl_device())
            t[self.rank] = mine
            dist.all_reduce(t, group=self.group)
            self._ntiles = [int(v) for v in t.tolist()]
        else:
            self._ntiles = [mine]

    # ---- stream plumbing: collectives are ordered against the backend's CUDA stream -----------------------------
    def _stream_ctx(self):
        if self.device.type == "cuda":
            return torch.cuda.stream(torch.cuda.ExternalStream(self.b.stream, device=self.device))
        import contextlib
        return contextlib.nullcontext()

    def _gather_async(self):
        if self.world == 1:
            return None
        return dist.all_gather_into_tensor(self.gathered, self.local, group=self.group, async_op=True)

    def _slot_ptr(self, r: int) -> int:
        return self.gathered.data_ptr() + r * self.slot_tiles * self.td * 8

    def _pairwise(self, pack, apply_first, apply_rest):
        """pack -> async all-gather -> own tiles -> wait -> every peer's tiles."""
        with self._stream_ctx():
            pack(self.local.data_ptr())
            work = self._gather_async()
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

    # ---- UJ_fmm over the sharded field: replicated tree, leaves split over the ranks ---------------------------------
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
        if self.world > 1:
            dist.all_reduce(cnt, group=self.group)
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
                recv = torch.empty((self.world, 7, slot), dtype=torch.float64, device=dev)
                dist.all_gather_into_tensor(recv.view(-1), send.view(-1), group=self.group)
                o = 0
                for r, c in enumerate(counts):
                    G[0:7, o:o + c] = recv[r, :, :c]
                    o += c
            else:
                G[0:7, :n_loc] = send[:, :n_loc]
            b.fmm_global(G.data_ptr(), G.shape[1], ntot, self.rank, self.world, 0)
            if self.world > 1:
                dist.all_reduce(G[9:24], group=self.group)
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
                if self.world > 1:
                    dist.all_reduce(G[12:15], group=self.group)
                state[39:42, :n_loc] += G[12:15, mine]

    # ---- pfield.UJ(pfield; reset, reset_sfs, sfs) -------------------------------------------------------------------
    def uj(self, reset: bool = True, reset_sfs: bool = False, sfs: bool = False):
        b = self.b
        if b.get_schemes().uj == _E.UJ_IDS["fmm"]:
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
            self.uj(True, True, True)
            st(_E.STAGE_STORE_TEST)
            st(_E.STAGE_SCALE_SIGMA_DOMAIN)
            self.uj(True, True, True)
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
