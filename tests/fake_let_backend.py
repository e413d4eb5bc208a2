"""CPU stand-in for the `vpmb200_let_*` phases of `Engine`, used ONLY by the gloo tests of the local-essential-tree
orchestration in flowunsteady_b200.dist (`_uj_fmm_let`, `_let_halo_evaluate`).

What is under test is the HOST choreography: the cube all-reduce, the histogram all-reduce and the cut, the all-to-all of
particle rows, the skeleton / multipole / record exchanges of both modes (all-gather, demand-driven halo: request and reply
all-to-alls with their per-owner counts and orderings), far-field reuse between DynamicSFS's two evaluations and the inverse
all-to-alls.  The "tree" here is deliberately trivial — a root and one leaf per occupied histogram bin — and so is the far
field (a monopole about the bin centre for bins at Chebyshev distance >= FAR, U only); the near field is the oracle's exact
pair sum.  With FAR = None everything is near field and the sharded result must equal the single-process oracle; with FAR = 2
the multipole channel carries data and the result must equal the SAME stand-in on one rank (what the GPU tests assert of the
real engine).  Not a product path.
"""
import ctypes as C

import numpy as np

from flowunsteady_b200 import engine as E
from oracle import oracle as o
from tests.fake_backend import FakeBackend

CELL = 8        # doubles per skeleton cell: start, count, bin, cx, cy, cz, nchild, leaf ordinal
_CT = {np.dtype(np.float64): C.c_double, np.dtype(np.int32): C.c_int32, np.dtype(np.int64): C.c_int64, np.dtype(np.uint8): C.c_uint8}


def _arr(ptr, n, dtype=np.float64):
    dtype = np.dtype(dtype)
    if n <= 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array((_CT[dtype] * int(n)).from_address(int(ptr)))


class FakeLetBackend(FakeBackend):
    FAR = None          # Chebyshev bin distance from which the monopole stands in for the leaf (None: all near field)

    def __init__(self, P_local, schemes_engine, schemes_oracle, far=None):
        super().__init__(P_local, schemes_engine, schemes_oracle)
        self.FAR = far
        self._n_own = 0
        self.calls = []      # phase log (the tests check the halo mode really took the request / reply route)

    # ---- plumbing the product backend does with device pointers -------------------------------------------------------
    def host_view(self, ptr, n, dtype):
        import torch
        npdt = {torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64, torch.uint8: np.uint8}[dtype]
        return torch.from_numpy(_arr(ptr, n, npdt))

    def synchronize(self):
        pass

    # ---- phases 1-3: bounds, bins + histogram, cut, rows to their owners ----------------------------------------------
    def let_bounds(self):
        if self.np == 0:
            return [1e300] * 3 + [-1e300] * 3
        x = self.P[:, E.X:E.X + 3]
        return list(x.min(0)) + list(x.max(0))

    def _bins(self, x):
        nb = self._nb
        ijk = np.minimum(((x - self._x0) / self._side * nb).astype(np.int64), nb - 1)
        return (ijk[:, 0] * nb + ijk[:, 1]) * nb + ijk[:, 2]

    def let_keys(self, lohi, level):
        lo, hi = np.asarray(lohi[:3], dtype=float), np.asarray(lohi[3:], dtype=float)
        side = float((hi - lo).max())
        self._side = side * (1 + 1e-9) if side > 0 else 1.0
        self._x0 = 0.5 * (lo + hi) - 0.5 * self._side
        self._nb = 1 << int(level)
        self._hbin = self._bins(self.P[:, E.X:E.X + 3]) if self.np else np.zeros(0, dtype=np.int64)
        self._hist = np.bincount(self._hbin, minlength=self._nb ** 3).astype(np.int32)
        self._work = np.zeros(self._nb ** 3, dtype=np.int64)
        return self._hist.ctypes.data, 0

    def let_work(self):
        return self._work.ctypes.data

    def let_partition(self, nparts, part, use_work):
        self._G, self._rank = int(nparts), int(part)
        hist = self._hist.astype(np.int64)          # all-reduced in place by the driver: the GLOBAL histogram
        before = np.cumsum(hist) - hist
        owner = np.minimum(self._G - 1, before * self._G // max(int(hist.sum()), 1))   # monotone, bin-aligned
        dest = owner[self._hbin]
        self._send_order = np.argsort(dest, kind="stable")
        return [int((dest == k).sum()) for k in range(self._G)]

    def let_pack(self, ptr):
        n = self.np
        if n:
            _arr(ptr, n * 7).reshape(n, 7)[:] = self.P[self._send_order, 0:7]      # X, Gamma, sigma

    # ---- phase 4: the owner's tree (root + one leaf per occupied bin), "multipoles", records ---------------------------
    def let_build(self, rows_ptr, n_own, n_all, reuse):
        rows = _arr(rows_ptr, n_own * 7).reshape(n_own, 7).copy() if n_own else np.zeros((0, 7))
        if reuse:
            assert n_own == self._n_own
        else:
            self._n_own = n_own
            b = self._bins(rows[:, :3]) if n_own else np.zeros(0, dtype=np.int64)
            self._perm = np.argsort(b, kind="stable")
            ub, start, count = np.unique(b[self._perm], return_index=True, return_counts=True)
            nl = len(ub)
            self._nl, self._nc = nl, (nl + 1 if n_own else 0)
            cells = np.zeros((max(self._nc, 1), CELL))
            if n_own:
                nb = self._nb
                ijk = np.stack([ub // (nb * nb), (ub // nb) % nb, ub % nb], axis=1)
                cells[0] = [0, n_own, -1, 0, 0, 0, nl, -1]
                cells[1:, 0], cells[1:, 1], cells[1:, 2] = start, count, ub
                cells[1:, 3:6] = self._x0 + (ijk + 0.5) * (self._side / nb)
                cells[1:, 7] = np.arange(nl)
            self._cells = cells
        srt = rows[self._perm]
        self._rec = np.zeros((max(n_own, 1), 10))
        self._rec[:n_own, :7] = srt
        if not reuse:
            self._M = np.zeros((max(self._nc, 1), 3))
            for k in range(self._nl):
                s, c = int(self._cells[1 + k, 0]), int(self._cells[1 + k, 1])
                self._M[1 + k] = srt[s:s + c, 3:6].sum(0)
            if n_own:
                self._M[0] = srt[:, 3:6].sum(0)
        self.calls.append("build_reuse" if reuse else "build")
        return (self._nc, self._nl if n_own else 0, 3, n_own)

    def let_ptrs(self):
        return self._cells.ctypes.data, self._M.ctypes.data, self._rec.ctypes.data

    @staticmethod
    def let_cell_bytes():
        return CELL * 8

    # ---- phase 5: every rank's leaves as sources -------------------------------------------------------------------------
    def _attach(self, cells_ptr, slot_c, nc):
        G = len(nc)
        allc = _arr(cells_ptr, G * slot_c * CELL).reshape(G, slot_c, CELL)
        S = dict(rank=[], cell=[], start=[], count=[], bin=[], centre=[])
        for q in range(G):
            for c in range(1, int(nc[q])):
                row = allc[q, c]
                S["rank"].append(q); S["cell"].append(c); S["start"].append(int(row[0])); S["count"].append(int(row[1]))
                S["bin"].append(int(row[2])); S["centre"].append(row[3:6].copy())
        nb = self._nb
        b = np.asarray(S["bin"], dtype=np.int64)
        self._S = {k: np.asarray(v) for k, v in S.items()}
        self._S["ijk"] = np.stack([b // (nb * nb), (b // nb) % nb, b % nb], axis=1) if len(b) else np.zeros((0, 3), dtype=np.int64)
        self._S["centre"] = self._S["centre"].reshape(-1, 3)

    def let_attach_tree(self, cells_ptr, M_ptr, slot_c, nc, np_):
        self._attach(cells_ptr, slot_c, nc)
        G = len(nc)
        self._M_all = _arr(M_ptr, G * slot_c * 3).reshape(G, slot_c, 3).copy()
        self._halo = False
        self.calls.append("attach_tree")

    def let_attach_records(self, rec_ptr, slot_n, np_):
        G = len(np_)
        allr = _arr(rec_ptr, G * slot_n * 10).reshape(G, slot_n, 10)
        self._recs = {q: allr[q, :int(np_[q])].copy() for q in range(G) if q != self._rank}
        self.calls.append("attach_records")

    def let_attach_skeleton(self, cells_ptr, slot_c, nc, np_, nl):
        assert [int(v) - 1 if v else 0 for v in nc] == [int(v) for v in nl]
        self._attach(cells_ptr, slot_c, nc)
        self._halo = True
        self._M2, self._rec2 = np.zeros((0, 3)), np.zeros((0, 10))
        self.calls.append("attach_skeleton")

    def _get_M(self, s):
        q, c = int(self._S["rank"][s]), int(self._S["cell"][s])
        if q == self._rank:
            return self._M[c]
        return self._M2[self._mslot[(q, c)]] if self._halo else self._M_all[q, c]

    def _get_rec(self, s):
        q, st, cn = int(self._S["rank"][s]), int(self._S["start"][s]), int(self._S["count"][s])
        if q == self._rank:
            return self._rec[st:st + cn]
        if self._halo:
            return self._rec2[self._poff[s]:self._poff[s] + cn]
        return self._recs[q][st:st + cn]

    # ---- phase 6: lists, far field, near field ---------------------------------------------------------------------------
    def _own_leaves(self):
        return [(int(r[0]), int(r[1]), int(r[2])) for r in self._cells[1:self._nc]]

    def let_evaluate(self, out_ptr, reuse, stage):
        n = self._n_own
        if n <= 0:
            return
        nb = self._nb
        if stage in (0, 1, 5) and not reuse:
            self._near, self._far = [], []
            for (st, cn, b) in self._own_leaves():
                t = np.array([b // (nb * nb), (b // nb) % nb, b % nb])
                d = np.abs(self._S["ijk"] - t).max(1)
                far = d >= self.FAR if self.FAR else np.zeros(len(d), dtype=bool)
                self._far.append(np.nonzero(far)[0])
                self._near.append(np.nonzero(~far)[0])
            self.calls.append("lists")
        if stage == 5:
            return
        if stage in (0, 1, 6) and not reuse:
            self._farU = np.zeros((n, 3))
            for k, (st, cn, b) in enumerate(self._own_leaves()):
                x = self._rec[st:st + cn, :3]
                for s in self._far[k]:
                    r = x - self._S["centre"][s]
                    self._farU[st:st + cn] += np.cross(self._get_M(s), r) / (4 * np.pi * (r * r).sum(1) ** 1.5)[:, None]
            self.calls.append("far")
        if stage in (1, 6):
            return
        self._sU, self._sJ = np.zeros((n, 3)), np.zeros((n, 9))
        for k, (st, cn, b) in enumerate(self._own_leaves()):
            src = np.concatenate([self._get_rec(s) for s in self._near[k]])
            U, J = o.uj_direct(self.kernel, src[:, 0:3], src[:, 3:6], src[:, 6], self._rec[st:st + cn, :3], accum=0)
            self._sU[st:st + cn] = U + self._farU[st:st + cn]
            self._sJ[st:st + cn] = J
        out = _arr(out_ptr, n * 12).reshape(n, 12)
        out[self._perm, 0:3] = self._sU
        out[self._perm, 3:12] = self._sJ
        self.calls.append("near")

    def let_estr_records(self):
        n = self._n_own
        if n <= 0:
            return
        Jm = self._sJ.reshape(-1, 3, 3).transpose(0, 2, 1)
        G = self._rec[:n, 3:6]
        self._rec[:n, 7:10] = np.einsum("plk,pl->pk", Jm, G) if self.so.transposed else np.einsum("pkl,pl->pk", Jm, G)

    def let_estr_evaluate(self, out_ptr):
        n = self._n_own
        if n <= 0:
            return
        sE = np.zeros((n, 3))
        zeta = np.vectorize(lambda q: o.zeta(self.kernel, q))
        for k, (st, cn, b) in enumerate(self._own_leaves()):
            src = np.concatenate([self._get_rec(s) for s in self._near[k]])
            x = self._rec[st:st + cn, :3]
            d = x[:, None, :] - src[None, :, 0:3]
            r = np.sqrt((d * d).sum(-1))
            z = zeta(r / src[None, :, 6]) / src[None, :, 6] ** 3
            Jm = self._sJ[st:st + cn].reshape(-1, 3, 3).transpose(0, 2, 1)
            a, bq = z @ src[:, 3:6], z @ src[:, 7:10]
            S = np.einsum("plk,pl->pk", Jm, a) if self.so.transposed else np.einsum("pkl,pl->pk", Jm, a)
            sE[st:st + cn] = S - bq
        _arr(out_ptr, n * 3).reshape(n, 3)[self._perm] = sE
        self.calls.append("estr")

    # ---- demand-driven halo ------------------------------------------------------------------------------------------------
    def let_halo_plan(self, nparts):
        want = [0] * (3 * nparts)
        far = sorted({int(s) for f in self._far for s in f if self._S["rank"][s] != self._rank}) if self._n_own > 0 else []
        near = sorted({int(s) for f in self._near for s in f if self._S["rank"][s] != self._rank}) if self._n_own > 0 else []
        # self._S is ordered by (rank, cell), so the sorted source indices are already grouped by owner in rank order
        self._mslot, self._poff = {}, {}
        rc, rl, off = [], [], 0
        for s in far:
            q = int(self._S["rank"][s])
            self._mslot[(q, int(self._S["cell"][s]))] = len(rc)
            rc.append(int(self._S["cell"][s]))
            want[3 * q] += 1
        for s in near:
            q, cn = int(self._S["rank"][s]), int(self._S["count"][s])
            self._poff[s] = off
            off += cn
            rl.append((int(self._S["start"][s]), cn))
            want[3 * q + 1] += 1
            want[3 * q + 2] += cn
        self._req_c = np.asarray(rc + [0], dtype=np.int32)
        self._req_l = np.asarray(rl + [(0, 0)], dtype=np.int32).reshape(-1, 2)
        self._want_c, self._want_r = len(rc), off
        self.calls.append("halo_plan")
        return want, self._req_c.ctypes.data, self._req_l.ctypes.data

    def let_halo_serve(self, req_cells_ptr, ncell, req_leaf_ptr, nleaf, M_out_ptr, rec_out_ptr):
        if M_out_ptr and ncell:
            ids = _arr(req_cells_ptr, ncell, np.int32)
            _arr(M_out_ptr, ncell * 3).reshape(ncell, 3)[:] = self._M[ids]
        if rec_out_ptr and nleaf:
            pairs = _arr(req_leaf_ptr, 2 * nleaf, np.int32).reshape(nleaf, 2)
            blk = np.concatenate([self._rec[s:s + c] for s, c in pairs])
            _arr(rec_out_ptr, blk.size).reshape(-1, 10)[:] = blk
        self.calls.append("halo_serve_M+rec" if M_out_ptr else "halo_serve_rec")

    def let_halo_set(self, M2_ptr, rec2_ptr):
        if M2_ptr and self._want_c:
            self._M2 = _arr(M2_ptr, self._want_c * 3).reshape(-1, 3).copy()
        if rec2_ptr and self._want_r:
            self._rec2 = _arr(rec2_ptr, self._want_r * 10).reshape(-1, 10).copy()

    # ---- phase 7: results back on the home rank ----------------------------------------------------------------------------
    def let_finish(self, res_ptr, which, reset):
        n = self.np
        if n == 0:
            return
        cols = 12 if which == 0 else 3
        res = _arr(res_ptr, n * cols).reshape(n, cols)
        idx = self._send_order
        if which == 0:
            if reset:
                self.P[idx, E.U:E.U + 3] = res[:, 0:3]
                self.P[idx, E.J:E.J + 9] = res[:, 3:12]
                self.P[idx, E.PSE:E.PSE + 3] = 0
            else:
                self.P[idx, E.U:E.U + 3] += res[:, 0:3]
                self.P[idx, E.J:E.J + 9] += res[:, 3:12]
        else:
            self.P[idx, E.SFS:E.SFS + 3] += res
