"""CPU stand-in for `Engine` used ONLY by the gloo tests of flowunsteady_b200.dist (host-side orchestration).

It speaks the same backend protocol (pack / from_records / stage) with numpy + the oracle's pair sums, so the
world_size-2 test can check partitioning, gather order, accumulate flags and stage sequencing against the
single-process oracle.  It is not a product path (the product backend is the CUDA engine).
"""
import ctypes as C

import numpy as np

from flowunsteady_b200 import engine as E
from oracle import oracle as o

TILE, REC, TD = 256, 10, 2570


def _view(ptr, ndoubles):
    return np.ctypeslib.as_array((C.c_double * ndoubles).from_address(ptr))


class FakeBackend:
    def __init__(self, P_local, schemes_engine, schemes_oracle):
        self.P = np.ascontiguousarray(P_local)
        self.se, self.so = schemes_engine, schemes_oracle
        self.t, self.nt = 0.0, 0
        self.stream = 0
        self.kernel = schemes_oracle.kernel

    np = property(lambda s: s.P.shape[0])

    @staticmethod
    def tiles_for(n):
        return (int(n) + TILE - 1) // TILE

    @staticmethod
    def tile_doubles():
        return TD

    def get_schemes(self):
        return self.se

    def get_time(self):
        return self.t, self.nt

    def set_time(self, t, nt):
        self.t, self.nt = t, nt

    def reset_particles(self):
        self.P[:, E.U:E.U + 3] = 0
        self.P[:, E.J:E.J + 9] = 0
        self.P[:, E.PSE:E.PSE + 3] = 0

    def reset_particles_sfs(self):
        self.P[:, E.SFS:E.SFS + 3] = 0

    # ---- tiles: records hold raw (x, Gamma, sigma[, v]); header slot 7 = number of real sources -------------------
    def _pack(self, ptr, extra=None):
        n, nt = self.np, self.tiles_for(self.np)
        buf = _view(ptr, nt * TD)
        buf[:] = 0
        for k in range(nt):
            lo, hi = k * TILE, min(n, (k + 1) * TILE)
            rec = buf[k * TD:k * TD + TILE * REC].reshape(TILE, REC)
            rec[:hi - lo, 0:3] = self.P[lo:hi, E.X:E.X + 3]
            rec[:hi - lo, 3:6] = self.P[lo:hi, E.GAMMA:E.GAMMA + 3]
            rec[:hi - lo, 6] = self.P[lo:hi, E.SIGMA]
            if extra is not None:
                rec[:hi - lo, 7:10] = extra[lo:hi]
            buf[k * TD + TILE * REC + 7] = hi - lo

    def _unpack(self, ptr, ntiles):
        buf = _view(ptr, ntiles * TD)
        rows = []
        for k in range(ntiles):
            nreal = int(buf[k * TD + TILE * REC + 7])
            rows.append(buf[k * TD:k * TD + TILE * REC].reshape(TILE, REC)[:nreal])
        return np.concatenate(rows) if rows else np.zeros((0, REC))

    def pack_uj_records(self, ptr):
        self._pack(ptr)

    def pack_estr_records(self, ptr):
        Jm = self.P[:, E.J:E.J + 9].reshape(-1, 3, 3).transpose(0, 2, 1)   # Jm[p, i, j] = J[i, j]
        G = self.P[:, E.GAMMA:E.GAMMA + 3]
        v = np.einsum("plk,pl->pk", Jm, G) if self.so.transposed else np.einsum("pkl,pl->pk", Jm, G)
        self._pack(ptr, extra=v)

    def uj_from_records(self, ptr, ntiles, accumulate):
        if not accumulate:
            self.P[:, E.U:E.U + 3] = 0
            self.P[:, E.J:E.J + 9] = 0
        if ntiles == 0 or self.np == 0:
            return
        src = self._unpack(ptr, ntiles)
        U, Jo = o.uj_direct(self.kernel, src[:, 0:3], src[:, 3:6], src[:, 6], self.P[:, E.X:E.X + 3], accum=1)
        self.P[:, E.U:E.U + 3] += U
        self.P[:, E.J:E.J + 9] += Jo

    def estr_from_records(self, ptr, ntiles):
        if ntiles == 0 or self.np == 0:
            return
        src = self._unpack(ptr, ntiles)
        x = self.P[:, E.X:E.X + 3]
        d = x[:, None, :] - src[None, :, 0:3]
        r = np.sqrt((d * d).sum(-1))
        sig = src[:, 6]
        z = np.vectorize(lambda q: o.zeta(self.kernel, q))(r / sig[None, :]) / sig[None, :] ** 3
        Jm = self.P[:, E.J:E.J + 9].reshape(-1, 3, 3).transpose(0, 2, 1)
        a = z @ src[:, 3:6]
        bq = z @ src[:, 7:10]
        S = np.einsum("plk,pl->pk", Jm, a) if self.so.transposed else np.einsum("pkl,pl->pk", Jm, a)
        self.P[:, E.SFS:E.SFS + 3] += S - bq

    # ---- UJ_fmm over the sharded field (dist.ShardedField._uj_fmm): the stand-in "FMM" is the exact sum -------------------
    def state_tensor(self):
        import torch
        return torch.from_numpy(self.P.T)            # (43, n) view: writes go straight into self.P

    def fmm_global(self, G_ptr, ldg, ntot, part, nparts, pass_):
        """Same contract as vpmb200_fmm_global: G holds (X, Gamma, sigma) of ALL particles in rows 0..6; this rank fills the
        U (9..11), J (15..23) rows [pass 0] or the E_str rows (12..14) [pass 1, reading the all-reduced J] of ITS share of
        the particles and zeros everywhere else, so that one all-reduce over the ranks assembles the field."""
        from flowunsteady_b200.dist import partition
        G = _view(G_ptr, 24 * ldg).reshape(24, ldg)
        lo, hi = partition(ntot, nparts)[part]
        X, Gm, sg = G[0:3, :ntot].T.copy(), G[3:6, :ntot].T.copy(), G[6, :ntot].copy()
        if pass_ == 0:
            G[9:24, :] = 0.0
            if hi > lo:
                U, J = o.uj_direct(self.kernel, X, Gm, sg, X[lo:hi], accum=0)
                G[9:12, lo:hi] = U.T
                G[15:24, lo:hi] = J.T
        else:
            G[12:15, :] = 0.0
            if hi > lo:
                Jall = G[15:24, :ntot].T.copy()
                Es = o.estr_direct(self.kernel, int(self.so.transposed), X, Gm, sg, Jall, X[lo:hi], Jall[lo:hi], accum=0)
                G[12:15, lo:hi] = np.asarray(Es).T

    # ---- per-particle stages through the oracle's exported pieces -------------------------------------------------
    def stage(self, stage, a=0.0, b=0.0, dt=0.0, Uinf=None):
        P, so = self.P, self.so
        live = ~(P[:, E.STATIC] > 0)
        zeta0 = o.zeta(self.kernel, 0.0)
        G = P[:, E.GAMMA:E.GAMMA + 3]
        Jm = P[:, E.J:E.J + 9].reshape(-1, 3, 3).transpose(0, 2, 1)
        S = np.einsum("plk,pl->pk", Jm, G) if so.transposed else np.einsum("pkl,pl->pk", Jm, G)
        if stage == E.STAGE_SCALE_SIGMA_TEST:
            P[live, E.SIGMA] *= so.alpha
        elif stage == E.STAGE_SCALE_SIGMA_DOMAIN:
            P[live, E.SIGMA] /= so.alpha
        elif stage == E.STAGE_STORE_TEST:
            P[live, E.M:E.M + 9] = 0
            P[live, E.M:E.M + 3] = S[live]
            P[live, E.M + 3:E.M + 6] = P[live, E.SFS:E.SFS + 3]
        elif stage == E.STAGE_DYNAMIC_COEFF:
            for i in np.nonzero(live)[0]:
                p = P[i]
                M1 = p[E.M:E.M + 3] - S[i]
                M2 = p[E.M + 3:E.M + 6] - p[E.SFS:E.SFS + 3]
                nume = (M1 @ G[i]) * (3 * so.alpha - 2)
                deno = (M2 @ G[i]) / (zeta0 / p[E.SIGMA] ** 3)
                if p[E.CC + 2] == 0:
                    p[E.CC + 2] = deno if deno != 0 else np.finfo(float).eps
                nume = so.sfs_rlxf * nume + (1 - so.sfs_rlxf) * p[E.CC + 1]
                deno = so.sfs_rlxf * deno + (1 - so.sfs_rlxf) * p[E.CC + 2]
                if abs(nume / deno) > so.maxC:
                    if abs(deno) < abs(p[E.CC + 2]):
                        deno = np.sign(deno) * abs(p[E.CC + 2])
                    nume = np.sign(nume) * abs(deno) * so.maxC
                elif abs(nume / deno) < so.minC:
                    nume = np.sign(nume) * abs(deno) * so.minC
                p[E.CC + 1], p[E.CC + 2] = nume, deno
                p[E.CC] = abs(nume / deno) if so.force_positive else nume / deno
                p[E.M:E.M + 9] = 0
        elif stage == E.STAGE_CONSTANT_COEFF:
            P[live, E.CC] = so.Cs
        elif stage == E.STAGE_CLIP_CONTROL:
            if so.clippings & 1:
                bad = live & (P[:, E.CC] * np.einsum("ij,ij->i", G, P[:, E.SFS:E.SFS + 3]) < 0)
                P[bad, E.CC] = 0
            assert so.controls == 0, "fake backend: controls not modelled"
        elif stage == E.STAGE_ZERO_M:
            P[live, E.M:E.M + 9] = 0
        elif stage in (E.STAGE_UPDATE, E.STAGE_UPDATE_EULER_RELAX):
            L = o.lib()
            u = np.ascontiguousarray(Uinf, dtype=np.float64)
            euler = so.integration == 0
            for i in np.nonzero(live)[0]:
                col = np.ascontiguousarray(P[i])
                keepM = col[E.M:E.M + 9].copy()
                if euler:
                    col[E.M:E.M + 9] = 0
                    sv = o.Schemes.from_buffer_copy(so)
                    sv.viscous = 0
                    L.vpmo_update_particle(col, C.byref(sv), 0.0, 1.0, dt, u, zeta0)
                    if stage == E.STAGE_UPDATE_EULER_RELAX:
                        L.vpmo_relax_particle(col, so.relaxation, so.rlxf)
                    if so.viscous == 1:
                        col[E.SIGMA] = np.sqrt(col[E.SIGMA] ** 2 + 2 * so.nu * dt)
                    col[E.M:E.M + 9] = keepM
                else:
                    L.vpmo_update_particle(col, C.byref(so), a, b, dt, u, zeta0)
                P[i] = col
        elif stage == E.STAGE_RELAX:
            L = o.lib()
            for i in np.nonzero(live)[0]:
                col = np.ascontiguousarray(P[i])
                L.vpmo_relax_particle(col, so.relaxation, so.rlxf)
                P[i] = col
        else:
            raise ValueError(stage)
