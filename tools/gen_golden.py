#!/usr/bin/env python
"""Generates tests/golden/uj_mp.npz: 50-digit mpmath evaluations of U, J = grad U and E_str for small
particle sets — the substitute pin for the (unpinned) hot path.

Independent of oracle/ and of the CUDA code:
  * U is the regularised Biot-Savart sum of docs/src/theory/rvpm.md:89-100 with the kernel form and
    the source-sigma convention of src/FLOWUnsteady_processing_force.jl:889-905;
  * J is obtained by NUMERICAL differentiation of that U (mpmath.diff at 50 digits), so it checks the
    analytic gradient expression instead of restating it;
  * E_str follows rvpm.md:251-263 with the index form of SURVEY.md A.4, using the mp J.

Run:  python tools/gen_golden.py     (takes a few minutes; output is committed)
"""
import os
import sys

import mpmath as mp
import numpy as np

mp.mp.dps = 50
FOURPI = 4 * mp.pi


def g_of(kernel, s):
    if kernel == "gaussianerf":
        return mp.erf(s / mp.sqrt(2)) - mp.sqrt(2 / mp.pi) * s * mp.exp(-s * s / 2)
    if kernel == "winckelmans":
        return s**3 * (s * s + mp.mpf("2.5")) / (s * s + 1) ** mp.mpf("2.5")
    if kernel == "gaussian":
        return 1 - mp.exp(-(s**3))
    if kernel == "singular":
        return mp.mpf(1)
    raise ValueError(kernel)


def zeta_of(kernel, s):
    if kernel == "gaussianerf":
        return mp.exp(-s * s / 2) / (2 * mp.pi) ** mp.mpf("1.5")
    if kernel == "winckelmans":
        return mp.mpf("7.5") / FOURPI / (s * s + 1) ** mp.mpf("3.5")
    if kernel == "gaussian":
        return 3 / FOURPI * mp.exp(-(s**3))
    raise ValueError(kernel)


def U_at(kernel, x, xs, gs, sig):
    u = [mp.mpf(0)] * 3
    for j in range(len(sig)):
        d = [x[k] - xs[j][k] for k in range(3)]
        r = mp.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
        if r == 0:
            continue
        c = -g_of(kernel, r / sig[j]) / (FOURPI * r**3)
        G = gs[j]
        u[0] += c * (d[1] * G[2] - d[2] * G[1])
        u[1] += c * (d[2] * G[0] - d[0] * G[2])
        u[2] += c * (d[0] * G[1] - d[1] * G[0])
    return u


def J_at(kernel, x, xs, gs, sig, skip):
    """J[i][j] = dU_i/dx_j by numerical differentiation; `skip` = index of the coincident source
    (its contribution is excluded, exactly as the r != 0 test excludes it at the particle itself)."""
    keep = [j for j in range(len(sig)) if j != skip]
    xs2 = [xs[j] for j in keep]
    gs2 = [gs[j] for j in keep]
    sg2 = [sig[j] for j in keep]
    Jm = [[None] * 3 for _ in range(3)]
    for jdir in range(3):
        for i in range(3):
            def f(h, i=i, jdir=jdir):
                xx = list(x)
                xx[jdir] = xx[jdir] + h
                return U_at(kernel, xx, xs2, gs2, sg2)[i]
            Jm[i][jdir] = mp.diff(f, 0, h=mp.mpf(10) ** (-20))
    return Jm


def main():
    out = {}
    rng = np.random.default_rng(20261017)
    n, nprobe = 40, 8
    x = rng.random((n, 3))
    gam = rng.standard_normal((n, 3)) / n
    sig = 2.125 * n ** (-1.0 / 3.0) * (0.6 + 0.8 * rng.random(n))
    # a few pathological pairs: nearly coincident, exactly coincident, and very far
    x[1] = x[0] + 1e-7
    x[3] = x[2]
    x[5] = x[4] + 40.0
    probes = rng.random((nprobe, 3)) * 1.5 - 0.25
    out["x"], out["gamma"], out["sigma"], out["probes"] = x, gam, sig, probes

    xs = [[mp.mpf(float(v)) for v in row] for row in x]
    gs = [[mp.mpf(float(v)) for v in row] for row in gam]
    sg = [mp.mpf(float(v)) for v in sig]
    targets = xs + [[mp.mpf(float(v)) for v in row] for row in probes]

    for kernel in ("gaussianerf", "winckelmans", "gaussian", "singular"):
        Uo = np.zeros((n + nprobe, 3))
        Jo = np.zeros((n + nprobe, 9))
        Jmp = []
        for i, xt in enumerate(targets):
            u = U_at(kernel, xt, xs, gs, sg)
            # coincident sources (r == 0) are skipped by U_at; for J exclude them explicitly
            coincident = [j for j in range(n) if all(xt[k] == xs[j][k] for k in range(3))]
            keep = [j for j in range(n) if j not in coincident]
            Jm = J_at(kernel, xt, [xs[j] for j in keep], [gs[j] for j in keep], [sg[j] for j in keep], skip=-1)
            Jmp.append(Jm)
            Uo[i] = [float(v) for v in u]
            for a in range(3):
                for b in range(3):
                    Jo[i, a + 3 * b] = float(Jm[a][b])
            print(kernel, i, file=sys.stderr)
        out[f"U_{kernel}"] = Uo
        out[f"J_{kernel}"] = Jo
        if kernel == "singular":
            continue
        # E_str at the particles (targets = sources), both stretching schemes, with the mp J
        for transposed in (1, 0):
            E = np.zeros((n, 3))
            for p in range(n):
                e = [mp.mpf(0)] * 3
                for q in range(n):
                    d = [xs[p][k] - xs[q][k] for k in range(3)]
                    r = mp.sqrt(d[0] ** 2 + d[1] ** 2 + d[2] ** 2)
                    z = zeta_of(kernel, r / sg[q]) / sg[q] ** 3
                    for k in range(3):
                        if transposed:
                            S = sum((Jmp[p][l][k] - Jmp[q][l][k]) * gs[q][l] for l in range(3))
                        else:
                            S = sum((Jmp[p][k][l] - Jmp[q][k][l]) * gs[q][l] for l in range(3))
                        e[k] += z * S
                E[p] = [float(v) for v in e]
            out[f"E_{kernel}_{'T' if transposed else 'C'}"] = E

    # kernel-function table: g, g', zeta at a sweep of r_hat (incl. tiny and large arguments)
    rh = np.concatenate([np.array([1e-8, 1e-5, 1e-3, 1e-2]), np.linspace(0.05, 12.0, 120), np.array([15.0, 30.0, 100.0])])
    out["rhat"] = rh
    for kernel in ("gaussianerf", "winckelmans", "gaussian"):
        gg = np.zeros_like(rh)
        dg = np.zeros_like(rh)
        zz = np.zeros_like(rh)
        for i, r in enumerate(rh):
            s = mp.mpf(float(r))
            gg[i] = float(g_of(kernel, s))
            dg[i] = float(mp.diff(lambda t: g_of(kernel, t), s))
            zz[i] = float(zeta_of(kernel, s))
        out[f"g_{kernel}"], out[f"dg_{kernel}"], out[f"zeta_{kernel}"] = gg, dg, zz

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "uj_mp.npz")
    np.savez_compressed(path, **out)
    print("wrote", os.path.normpath(path))


if __name__ == "__main__":
    main()
