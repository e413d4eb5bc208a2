#!/usr/bin/env python
"""Generates flowunsteady_b200/csrc/fmm_ops.inc — straight-line Cartesian-Taylor FMM operators for the Laplace
Green's function 1/r, specialised per expansion order P (vpm.FMM(; p), /root/reference/src/FLOWUnsteady_simulation.jl:43).

Conventions.  Multi-indices a = (i, j, k), graded enumeration idx(a).
    PM = P - 1   highest multipole order            (ExaFMM's P counts terms n = 0 .. P-1)
    PL = P + 1   highest local order                (U needs one derivative of the potential, J two: the two extra
                                                     orders keep J at the multipole truncation order)
    Mt_a = sum_s q_s (-1)^|a| (x_s - c)^a / a!        normalised multipole about c
    D_g  = d^g (1/r) at R = c_target - c_source       derivative tensor, |g| <= PL
    Lt_b = sum_{|a| <= min(PM, PL-|b|)} D_{a+b} Mt_a  local coefficients:  phi(c + u) = sum_b Lt_b u^b / b!
All operators are pure FMA chains with compile-time indices (no tables at run time).
"""
import itertools
import math
import os

ORDERS = (2, 3, 4, 5, 6)


def multi_indices(nmax):
    out = []
    for n in range(nmax + 1):
        for i in range(n, -1, -1):
            for j in range(n - i, -1, -1):
                out.append((i, j, n - i - j))
    return out


def fact(a):
    return math.factorial(a[0]) * math.factorial(a[1]) * math.factorial(a[2])


def order(a):
    return sum(a)


def add(a, b):
    return (a[0] + b[0], a[1] + b[1], a[2] + b[2])


def sub(a, b):
    return (a[0] - b[0], a[1] - b[1], a[2] - b[2])


def leq(a, b):
    return all(x <= y for x, y in zip(a, b))


def nterms(n):
    return (n + 1) * (n + 2) * (n + 3) // 6


def mono_code(var, nmax, idx, name, scale_fact=True, sign=False):
    """Emit code computing name[idx(a)] = (sign ? (-1)^|a| : 1) * var^a / a!  for |a| <= nmax (recursively)."""
    lines = [f"    double {name}[{nterms(nmax)}];", f"    {name}[0] = 1.0;"]
    comp = ("x", "y", "z")
    for a in multi_indices(nmax)[1:]:
        # pick the first nonzero component to peel off
        for c in range(3):
            if a[c] > 0:
                e = [0, 0, 0]
                e[c] = 1
                prev = sub(a, tuple(e))
                coef = 1.0 / a[c] if scale_fact else 1.0
                if sign:
                    coef = -coef
                lines.append(f"    {name}[{idx[a]}] = {name}[{idx[prev]}] * ({var}{comp[c]} * {coef!r});")
                break
    return lines


def gen_order(P):
    PM, PL = P - 1, P + 1
    idx = {a: n for n, a in enumerate(multi_indices(PL))}
    NM, NL = nterms(PM), nterms(PL)
    out = []
    out.append(f"// ---------------------------------------------------------------- P = {P}: PM = {PM} ({NM} terms), PL = {PL} ({NL} terms)")
    out.append(f"template <> struct FmmOps<{P}> {{")
    out.append(f"    static constexpr int PM = {PM}, PL = {PL}, NM = {NM}, NL = {NL};")

    # ---- derivative tensor D_g = g! b_g, b from the Duan-Krasny recurrence
    out.append("    // D[idx(g)] = d^g (1/r) at (x, y, z), |g| <= PL")
    out.append("    __device__ __forceinline__ static void dtensor(double x, double y, double z, double* __restrict__ D) {")
    out.append("        const double r2 = x * x + y * y + z * z;")
    out.append("        const double ir2 = 1.0 / r2;")
    out.append(f"        double b[{NL}];")
    out.append("        b[0] = rsqrt(r2);")
    comp = ("x", "y", "z")
    for g in multi_indices(PL)[1:]:
        n = order(g)
        t1 = []
        t2 = []
        for c in range(3):
            e = [0, 0, 0]
            e[c] = 1
            if g[c] >= 1:
                t1.append(f"{comp[c]} * b[{idx[sub(g, tuple(e))]}]")
            if g[c] >= 2:
                e[c] = 2
                t2.append(f"b[{idx[sub(g, tuple(e))]}]")
        expr = f"{(2 * n - 1) / n!r} * ({' + '.join(t1)})"
        if t2 and n > 1:
            expr += f" + {(n - 1) / n!r} * ({' + '.join(t2)})"
        out.append(f"        b[{idx[g]}] = -ir2 * ({expr});")
    for g in multi_indices(PL):
        out.append(f"        D[{idx[g]}] = b[{idx[g]}] * {float(fact(g))!r};")
    out.append("    }")

    # ---- M2L
    npair = 0
    out.append("    // L[b * LS] += sum_a D[a + b] * M[a]   (one scalar component; LS = stride of the accumulator array)")
    out.append("    template <int LS>")
    out.append("    __device__ __forceinline__ static void m2l(const double* __restrict__ D, const double* __restrict__ M, double* __restrict__ L) {")
    for b in multi_indices(PL):
        terms = []
        for a in multi_indices(min(PM, PL - order(b))):
            terms.append((idx[add(a, b)], idx[a]))
        npair += len(terms)
        expr = f"L[{idx[b]} * LS]"
        out.append(f"        {{ double acc = {expr};")
        for (dg, ma) in terms:
            out.append(f"          acc = fma(D[{dg}], M[{ma}], acc);")
        out.append(f"          L[{idx[b]} * LS] = acc; }}")
    out.append("    }")
    out.append(f"    static constexpr int M2L_FMAS = {npair};")

    # ---- P2M:  M[a] += q * (-1)^|a| v^a / a!
    out.append("    // M[a] += q_c * (-v)^a / a!  for the three components; v = x_s - c")
    out.append("    __device__ __forceinline__ static void p2m(double vx, double vy, double vz, double q0, double q1, double q2,")
    out.append("                                               double* __restrict__ M0, double* __restrict__ M1, double* __restrict__ M2) {")
    out += ["    " + l for l in mono_code("v", PM, idx, "w", scale_fact=True, sign=True)]
    out.append(f"        for (int a = 0; a < {NM}; ++a) {{ M0[a] = fma(q0, w[a], M0[a]); M1[a] = fma(q1, w[a], M1[a]); M2[a] = fma(q2, w[a], M2[a]); }}")
    out.append("    }")

    # ---- M2M:  Mp[a] += sum_{b <= a} Mc[b] * (-d)^(a-b)/(a-b)!,  d = c_child - c_parent
    out.append("    // parent += shift(child);  d = c_child - c_parent")
    out.append("    __device__ __forceinline__ static void m2m(double dx, double dy, double dz, const double* __restrict__ Mc, double* __restrict__ Mp) {")
    out += ["    " + l for l in mono_code("d", PM, idx, "w", scale_fact=True, sign=True)]
    for a in multi_indices(PM):
        terms = [f"Mc[{idx[b]}] * w[{idx[sub(a, b)]}]" for b in multi_indices(PM) if leq(b, a)]
        out.append(f"        Mp[{idx[a]}] += {' + '.join(terms)};")
    out.append("    }")

    # ---- L2L:  Lc[k] += sum_{b >= k} Lp[b] * e^(b-k)/(b-k)!,  e = c_child - c_parent
    out.append("    // child += shift(parent);  e = c_child - c_parent")
    out.append("    __device__ __forceinline__ static void l2l(double ex, double ey, double ez, const double* __restrict__ Lp, double* __restrict__ Lc) {")
    out += ["    " + l for l in mono_code("e", PL, idx, "w", scale_fact=True, sign=False)]
    for k in multi_indices(PL):
        terms = [f"Lp[{idx[b]}] * w[{idx[sub(b, k)]}]" for b in multi_indices(PL) if leq(k, b)]
        out.append(f"        Lc[{idx[k]}] += {' + '.join(terms)};")
    out.append("    }")

    # ---- L2P: gradient (3) and Hessian (6: xx, xy, xz, yy, yz, zz) of phi at c + u
    out.append("    // g[i] = d_i phi, h = (xx, xy, xz, yy, yz, zz) second derivatives at c + u, from the local expansion L")
    out.append("    __device__ __forceinline__ static void l2p(double ux, double uy, double uz, const double* __restrict__ L, double* __restrict__ g, double* __restrict__ h) {")
    out += ["    " + l for l in mono_code("u", PL - 1, idx, "w", scale_fact=True, sign=False)]
    E = [(1, 0, 0), (0, 1, 0), (0, 0, 1)]
    for i in range(3):
        terms = [f"L[{idx[add(b, E[i])]}] * w[{idx[b]}]" for b in multi_indices(PL - 1)]
        out.append(f"        g[{i}] = {' + '.join(terms)};")
    hp = [(0, 0), (0, 1), (0, 2), (1, 1), (1, 2), (2, 2)]
    for n, (i, j) in enumerate(hp):
        terms = [f"L[{idx[add(add(b, E[i]), E[j])]}] * w[{idx[b]}]" for b in multi_indices(PL - 2)]
        out.append(f"        h[{n}] = {' + '.join(terms)};")
    out.append("    }")
    out.append("};")
    return out


def main():
    lines = ["// GENERATED by tools/gen_fmm_ops.py — do not edit.", "#pragma once", "",
             "template <int P> struct FmmOps;", ""]
    for P in ORDERS:
        lines += gen_order(P)
        lines.append("")
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "flowunsteady_b200", "csrc", "fmm_ops.inc")
    open(path, "w").write("\n".join(lines))
    print("wrote", os.path.normpath(path), len(lines), "lines")


if __name__ == "__main__":
    main()
