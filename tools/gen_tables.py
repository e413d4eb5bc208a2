#!/usr/bin/env python
"""Generates flowunsteady_b200/csrc/gauss_table.inc — piecewise-polynomial tables for the
Gaussian-erf regularising function, in the variable t = (r/sigma)^2 (so the near-field needs neither
sqrt nor erf nor exp nor a division):

    G(t) = g(s)/s^3                      ->  g/r^3                       = G(t)/sigma^3
    H(t) = (g'(s)/s^2 - 3 G(t))/t        ->  (g'/(sigma r) - 3 g/r^2)/r^3 = H(t)/sigma^5
    Z(t) = exp(-t/2)                     ->  zeta(s) = (2 pi)^(-3/2) Z(t)

with g, g', zeta of SURVEY.md A.3 (gaussianerf, the reference default at
src/FLOWUnsteady_simulation.jl:37).  All three are entire functions of t.  Intervals are uniform,
centred at i*W (i = rint(t/W)); coefficients are monomials in u = t - i*W, fitted at Chebyshev nodes in
60-digit arithmetic.  Beyond T_FAR the kernel is singular to double precision (|1 - g| < 2^-56).

The pair kernels are bound by shared-memory bandwidth on the table lookups as much as by the FP64 pipe, so the FP64
tables are laid out for the fewest bytes per lookup:
  * H = 2 dG/dt identically (differentiate g/s^3 with respect to t = s^2), so ONE polynomial serves both: the kernel
    evaluates G and its derivative in the same Horner recurrence (value + derivative = 17 FMAs) from 10 coefficients
    = 80 B per lookup (five LDS.128) instead of 16 coefficients = 128 B (eight).  Width 1/2 with degree 9 keeps both to
    3e-16.  (A width-3/16 degree-7 variant reads 64 B per lookup but the lanes of a warp then fall into ~3x more distinct
    intervals, which costs more shared-memory wavefronts than it saves: measured 842 M "bank conflicts" per 1M-particle
    near-field pass, profiles/r01b_fmm_leaf_n1m.txt.)
  * Z(t) = exp(-c_i/2) * exp(-u/2): one tabulated double per interval (width 1/8), the second factor is a degree-7
    Taylor polynomial with constant coefficients (|u/2| <= 1/32: truncation 2e-17) — 8 B per lookup instead of 64 B.

Run:  python tools/gen_tables.py [--check]
"""
import argparse
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 60
W = mp.mpf("0.5")        # interval width of the FP32 tables
DEG = 7                  # polynomial degree (FP64 tables)
WG = mp.mpf("0.5")       # FP64 G table: interval width (exact in binary), value + derivative from one polynomial
GDEG = 9                 # ... of this degree (odd: coefficients are stored in pairs)
WZ = mp.mpf("0.125")     # FP64 Z table: interval width; Z = E[i] * taylor(exp(-u/2))
T_FAR = 88.0             # |1 - g(sqrt(t))| < 1.2e-17 beyond this
# FP32 variant of the pair kernel: lower degree, shorter range (|1 - g| < 3e-8 beyond T_FAR32)
DEG32 = 4
T_FAR32 = 40.0

SQ2PI = mp.sqrt(2 / mp.pi)


def G(t):
    t = mp.mpf(t)
    if t < mp.mpf("1e-6"):
        # series: sqrt(2/pi) sum_{n>=1} (-1)^(n+1) t^(n-1) 2n / ((2n+1) 2^n n!)
        return SQ2PI * sum((-1) ** (n + 1) * t ** (n - 1) * 2 * n / ((2 * n + 1) * 2**n * mp.factorial(n))
                           for n in range(1, 12))
    s = mp.sqrt(t)
    return (mp.erf(s / mp.sqrt(2)) - SQ2PI * s * mp.exp(-t / 2)) / (t * s)


def H(t):
    t = mp.mpf(t)
    if t < mp.mpf("1e-6"):
        # g'/s^2 = sqrt(2/pi) sum_n (-1)^n t^n/(2^n n!) ;  H = sum_n t^(n-1) [e_n - 3 G_n]
        tot = mp.mpf(0)
        for n in range(1, 12):
            e_n = (-1) ** n / (2**n * mp.factorial(n))
            g_n = (-1) ** n * 2 * (n + 1) / ((2 * n + 3) * 2 ** (n + 1) * mp.factorial(n + 1))  # coeff of t^n in G/sqrt(2/pi)
            tot += t ** (n - 1) * (e_n - 3 * g_n)
        return SQ2PI * tot
    return (SQ2PI * mp.exp(-t / 2) - 3 * G(t)) / t


def Z(t):
    return mp.exp(-mp.mpf(t) / 2)


def fit(func, c, half, deg):
    """Monomial coefficients (in u = t - c) of the degree-`deg` interpolant at Chebyshev nodes."""
    n = deg + 1
    nodes = [mp.cos(mp.pi * (2 * k + 1) / (2 * n)) for k in range(n)]
    A = mp.matrix(n, n)
    b = mp.matrix(n, 1)
    for r, xk in enumerate(nodes):
        u = xk * half
        for k in range(n):
            A[r, k] = u**k
        b[r] = func(c + u)
    sol = mp.lu_solve(A, b)
    return [sol[k] for k in range(n)]


def horner64(coefs, u):
    acc = np.float64(coefs[-1])
    for ck in coefs[-2::-1]:
        acc = acc * u + np.float64(ck)   # numpy float64; an FMA on the device is at least as accurate
    return acc


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true", help="also measure the double-precision error")
    args = ap.parse_args()

    nint = int(mp.ceil(mp.mpf(T_FAR) / WG)) + 2   # centres 0 .. nint-1 (rint can round up at T_FAR)
    nintz = int(mp.ceil(mp.mpf(T_FAR) / WZ)) + 2
    half = WG / 2
    # the first interval only needs [0, W/2]; fit it on [-W/2, W/2] anyway (G is entire)
    tabG = [[float(v) for v in fit(G, i * WG, half, GDEG)] for i in range(nint)]
    tabE = [float(mp.exp(-(i * WZ) / 2)) for i in range(nintz)]
    zc = [float((-mp.mpf(1) / 2) ** k / mp.factorial(k)) for k in range(DEG + 1)]   # exp(-u/2) Taylor coefficients

    if args.check:
        rng = np.random.default_rng(1)
        worst = {"G": 0.0, "H": 0.0, "Z": 0.0}
        for i in range(nint - 1):
            c = float(i * WG)
            for t in np.concatenate([c + (rng.random(24) - 0.5) * float(WG), [c - 0.4999 * float(WG), c + 0.4999 * float(WG)]]):
                if t < 0:
                    continue
                u = np.float64(t) - np.float64(c)
                co = tabG[i]
                p, d = np.float64(co[-1]), np.float64(0.0)
                for ck in co[-2::-1]:
                    d = d * u + p
                    p = p * u + np.float64(ck)
                for name, got, ref in (("G", p, G(t)), ("H", 2 * d, H(t))):
                    worst[name] = max(worst[name], float(abs((mp.mpf(float(got)) - ref) / ref)))
        for t in rng.random(4000) * T_FAR:
            i = int(np.rint(t / float(WZ)))
            u = np.float64(t) - np.float64(i * float(WZ))
            got = horner64(zc, u) * np.float64(tabE[i])
            worst["Z"] = max(worst["Z"], float(abs((mp.mpf(float(got)) - Z(t)) / Z(t))))
        print("max relative error (float64 Horner):", worst)
        print("1 - g at T_FAR:", float(1 - G(T_FAR) * mp.mpf(T_FAR) ** mp.mpf("1.5")))

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flowunsteady_b200", "csrc",
                        "gauss_table.inc")
    with open(path, "w") as f:
        f.write("// GENERATED by tools/gen_tables.py — do not edit.\n#pragma once\n")
        f.write("// t = (r/sigma)^2.  G table: degree-%d polynomials in u = t - i*W, i = rint(t/W); entry [j][i] (a double2) =\n" % GDEG)
        f.write("// {c_2j, c_2j+1}, the coefficients of u^2j and u^(2j+1) on interval i.  G(t) = p(u), H(t) = 2 p'(u).\n")
        f.write("// Z table: Z(t) = vpm_gt_Z[i] * sum_k VPM_GZ_Ck u^k with i = rint(t/WZ), u = t - i*WZ.\n")
        f.write("#define VPM_GT_W %s\n" % repr(float(W)))
        f.write("#define VPM_GT_INVW %s\n" % repr(float(1 / W)))
        f.write("#define VPM_GT_DEG %d\n" % DEG)
        f.write("#define VPM_GT_TFAR %s\n" % repr(T_FAR))
        f.write("#define VPM_GG_W %s\n" % repr(float(WG)))
        f.write("#define VPM_GG_INVW %s\n" % repr(float(1 / WG)))
        f.write("#define VPM_GG_NINT %d\n" % nint)
        f.write("#define VPM_GG_DEG %d\n" % GDEG)
        f.write("#define VPM_GG_DOUBLES ((VPM_GG_DEG + 1) * VPM_GG_NINT)\n")
        f.write("#define VPM_GZ_W %s\n" % repr(float(WZ)))
        f.write("#define VPM_GZ_INVW %s\n" % repr(float(1 / WZ)))
        f.write("#define VPM_GZ_NINT %d\n" % nintz)
        for k in range(DEG + 1):
            f.write("#define VPM_GZ_C%d %s\n" % (k, repr(zc[k])))
        f.write("static const double vpm_gt_GH[VPM_GG_DOUBLES] = {\n")
        for j in range((GDEG + 1) // 2):
            for i in range(nint):
                f.write("  %s, %s,\n" % (repr(tabG[i][2 * j]), repr(tabG[i][2 * j + 1])))
        f.write("};\n")
        f.write("static const double vpm_gt_Z[VPM_GZ_NINT] = {\n")
        for i in range(nintz):
            f.write("  %s,\n" % repr(tabE[i]))
        f.write("};\n")
        # FP32 table (same interval width)
        nint32 = int(mp.ceil(mp.mpf(T_FAR32) / W)) + 2
        f.write("#define VPM_GT32_DEG %d\n" % DEG32)
        f.write("#define VPM_GT32_NINT %d\n" % nint32)
        f.write("#define VPM_GT32_TFAR %sf\n" % repr(T_FAR32))
        f.write("static const float vpm_gt32_GH[(VPM_GT32_DEG + 1) * VPM_GT32_NINT * 2] = {\n")
        t32 = [[fit(fn, i * W, half, DEG32) for fn in (G, H)] for i in range(nint32)]
        for k in range(DEG32 + 1):
            for i in range(nint32):
                f.write("  %sf, %sf,\n" % (repr(float(np.float32(float(t32[i][0][k])))), repr(float(np.float32(float(t32[i][1][k]))))))
        f.write("};\n")
    print("wrote", os.path.normpath(path), "intervals:", nint, "degree:", DEG)


if __name__ == "__main__":
    main()
