/*
 * fmm_oracle.c — CPU restatement of the REFERENCE's fast multipole evaluation of U and J = grad U.  TEST INFRASTRUCTURE ONLY.
 *
 * What it follows.  FLOWUnsteady selects `vpm_UJ = vpm.UJ_fmm` with `vpm_fmm = vpm.FMM(; p=4, ncrit=50, theta=0.4,
 * nonzero_sigma=false)` (/root/reference/src/FLOWUnsteady_simulation.jl:38,43,129) and documents the method as "the fast
 * multipole method ... approximating the velocity field and vortex stretching through SPHERICAL HARMONICS" in a modified
 * ExaFMM (/root/reference/docs/src/theory/rvpm.md:369-370; README.md:115).  ExaFMM itself is not in /root/reference (an
 * un-vendored dependency, SURVEY.md §2.2: "FLOWExaFMM.jl, C++ ExaFMM fork", no pinned commit), so this file restates its
 * PUBLISHED algorithm — Laplace kernel, solid-harmonic expansions truncated at degree p - 1 for multipoles AND locals,
 * adaptive octree with at most ncrit particles per leaf, dual tree traversal with the acceptance (R_i + R_j) < theta |c_i - c_j|
 * on the cells' half sides (Yokota & Barba, "ExaFMM"; Barba & Yokota JOSS 2021) — for the vector potential
 * psi = (1 / 4 pi) sum Gamma / r of the three strength components, U = curl psi and J = grad U taken from the derivatives of the
 * TRUNCATED local expansion (the reference obtains J by complex-step differentiation of the same expansion, rvpm.md:370 —
 * analytically identical), near field = the regularised pair kernel of vpm_oracle.c (vpmo_uj_direct).  PARITY UNPINNED: no
 * golden vector for the FMM exists in the reference; what this oracle pins is the error LEVEL of the method at a given (p,
 * theta, ncrit) against the direct sum, which tests/test_fmm_oracle.py and tests/test_gpu_fmm.py compare with the CUDA FMM
 * (Cartesian Taylor, multipoles to order p - 1, locals to order p + 1) at the same settings.
 *
 * Solid harmonics (normalisation of Dehnen 2014; identities verified numerically, tools/sh_identities.py):
 *   R_n^m(x) = r^n P_n^m(cos t) e^{i m phi} / (n + m)!          I_n^m(x) = (n - m)! P_n^m(cos t) e^{i m phi} / r^{n+1}
 *   1 / |x - y| = sum_{n,m} conj(R_n^m(y)) I_n^m(x)                                   (|y| < |x|)
 *   R_n^m(a + b) = sum_{k,l} R_k^l(a) R_{n-k}^{m-l}(b)
 *   I_n^m(a + b) = sum_{k,l} (-1)^k conj(R_k^l(a)) I_{n+k}^{m+l}(b)                    (|a| < |b|)
 *   d/dz R_n^m = R_{n-1}^m,  (d/dx + i d/dy) R_n^m = R_{n-1}^{m+1},  (d/dx - i d/dy) R_n^m = -R_{n-1}^{m-1}
 */
#include <complex.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "vpm_oracle.h"

static const double CONST4 = 0.07957747154594767;   /* 1 / (4 pi), the literal of vpm_oracle.c */
#define FO_MAXP 8          /* expansion order p: degrees 0 .. p - 1 */
#define FO_MAXLEVEL 21
#define FO_NCOEF (FO_MAXP * FO_MAXP)
typedef double complex cplx;

static inline int idx_nm(int n, int m) { return n * n + n + m; } /* m in -n .. n */

/* all R_n^m (regular) or I_n^m (irregular), 0 <= n <= N, -n <= m <= n */
static void harmonics(const double x[3], int N, int irregular, cplx *out) {
    const double X = x[0], Y = x[1], Z = x[2], r2 = X * X + Y * Y + Z * Z;
    const cplx xy = X + I * Y;
    for (int k = 0; k < (N + 1) * (N + 1); ++k) out[k] = 0;
    out[idx_nm(0, 0)] = irregular ? 1.0 / sqrt(r2) : 1.0;
    for (int m = 0; m <= N; ++m) {
        if (m > 0)
            out[idx_nm(m, m)] = irregular ? -(2.0 * m - 1.0) * xy / r2 * out[idx_nm(m - 1, m - 1)]
                                          : -xy / (2.0 * m) * out[idx_nm(m - 1, m - 1)];
        for (int n = m + 1; n <= N; ++n) {
            const cplx a = out[idx_nm(n - 1, m)], b = n - 2 >= m ? out[idx_nm(n - 2, m)] : 0;
            out[idx_nm(n, m)] = irregular ? ((2.0 * n - 1.0) * Z * a - (double)((n - 1) * (n - 1) - m * m) * b) / r2
                                          : ((2.0 * n - 1.0) * Z * a - r2 * b) / (double)((n + m) * (n - m));
        }
    }
    for (int n = 1; n <= N; ++n)
        for (int m = 1; m <= n; ++m) out[idx_nm(n, -m)] = ((m & 1) ? -1.0 : 1.0) * conj(out[idx_nm(n, m)]);
}

typedef struct {
    int start, count, parent, child0, nchild, level;
    double c[3], R;
} fo_cell;

typedef struct {
    int p, ncrit, n;
    double theta;
    const double *x, *g, *s;     /* Morton-ordered copies */
    fo_cell *cells;
    int ncells, cap;
    cplx *M, *L;                 /* [cell][3][p*p] */
    double *U, *J;               /* Morton order */
    int32_t kernel;
    int64_t n_m2l, n_p2p;
} fo_tree;

static uint64_t spread3(uint64_t v) {
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

typedef struct { uint64_t key; int idx; } fo_kv;
static int cmp_kv(const void *a, const void *b) {
    const fo_kv *p = a, *q = b;
    if (p->key != q->key) return p->key < q->key ? -1 : 1;
    return p->idx - q->idx;
}

static int lower_bound(const fo_kv *kv, int lo, int hi, uint64_t v) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (kv[mid].key < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

static void p2p(fo_tree *t, const fo_cell *ci, const fo_cell *cj) {
    vpmo_uj_direct(t->kernel, cj->count, t->x + 3 * cj->start, t->g + 3 * cj->start, t->s + cj->start, ci->count,
                   t->x + 3 * ci->start, t->U + 3 * ci->start, t->J + 9 * ci->start, 0);
    t->n_p2p++;
}

/* L_i[k,l] += (-1)^k sum_{n,m} M_j[n,m] I_{n+k}^{m+l}(c_i - c_j) */
static void m2l(fo_tree *t, int i, int j) {
    const int p = t->p, N = 2 * (p - 1);
    double D[3] = {t->cells[i].c[0] - t->cells[j].c[0], t->cells[i].c[1] - t->cells[j].c[1], t->cells[i].c[2] - t->cells[j].c[2]};
    cplx Ih[(2 * FO_MAXP - 1) * (2 * FO_MAXP - 1)];
    harmonics(D, N, 1, Ih);
    for (int comp = 0; comp < 3; ++comp) {
        const cplx *Mj = t->M + ((size_t)j * 3 + comp) * FO_NCOEF;
        cplx *Li = t->L + ((size_t)i * 3 + comp) * FO_NCOEF;
        for (int k = 0; k < p; ++k)
            for (int l = -k; l <= k; ++l) {
                cplx acc = 0;
                for (int n = 0; n < p; ++n)
                    for (int m = -n; m <= n; ++m)
                        if (abs(m + l) <= n + k) acc += Mj[idx_nm(n, m)] * Ih[idx_nm(n + k, m + l)];
                Li[idx_nm(k, l)] += (k & 1) ? -acc : acc;
            }
    }
    t->n_m2l++;
}

/* the traversal rule of the CUDA path (fmm.cuh: fmm_traverse_kernel) == ExaFMM's dual tree traversal */
static void traverse(fo_tree *t, int i, int j) {
    const fo_cell *ci = &t->cells[i], *cj = &t->cells[j];
    const double dx = ci->c[0] - cj->c[0], dy = ci->c[1] - cj->c[1], dz = ci->c[2] - cj->c[2];
    const double d2 = dx * dx + dy * dy + dz * dz, rs = ci->R + cj->R;
    if (rs * rs < t->theta * t->theta * d2) {
        m2l(t, i, j);
    } else if (ci->nchild == 0 && cj->nchild == 0) {
        p2p(t, ci, cj);
    } else if (cj->nchild == 0 || (ci->nchild != 0 && ci->R >= cj->R)) {
        for (int k = 0; k < ci->nchild; ++k) traverse(t, ci->child0 + k, j);
    } else {
        for (int k = 0; k < cj->nchild; ++k) traverse(t, i, cj->child0 + k);
    }
}

/*
 * U (nt x 3) and J (nt x 9, J[i + 3 j] = du_i/dx_j) at every particle from every particle, by the FMM described above.
 * kernel: VPMO_KERNEL_* of the NEAR field (the far field is singular: nonzero_sigma = false).  leaf_sigmas > 0 mirrors the CUDA
 * path's sparse-leaf refinement (a would-be leaf is split while its half side exceeds leaf_sigmas core sizes); 0 = plain ncrit.
 * stats[0..3] = cells, leaves, M2L pairs, P2P pairs.  Returns 0 on success.
 */
int32_t vpmo_fmm_uj(int32_t kernel, int32_t p, int32_t ncrit, double theta, double leaf_sigmas, int64_t n, const double *x,
                    const double *g, const double *sig, double *U, double *J, int64_t *stats) {
    if (p < 1 || p > FO_MAXP || ncrit < 1 || n <= 0 || !(theta > 0)) return -1;
    /* root cube and keys: the arithmetic of fmm_sort / fmm_keys_kernel */
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int64_t i = 0; i < n; ++i)
        for (int c = 0; c < 3; ++c) {
            lo[c] = fmin(lo[c], x[3 * i + c]);
            hi[c] = fmax(hi[c], x[3 * i + c]);
        }
    double side = fmax(fmax(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    side = side > 0 ? side * (1.0 + 1e-9) : 1.0;
    const double ctr[3] = {0.5 * (lo[0] + hi[0]), 0.5 * (lo[1] + hi[1]), 0.5 * (lo[2] + hi[2])};
    const double x0[3] = {ctr[0] - 0.5 * side, ctr[1] - 0.5 * side, ctr[2] - 0.5 * side}, inv = 2097152.0 / side;
    fo_kv *kv = malloc(sizeof(fo_kv) * n);
    for (int64_t i = 0; i < n; ++i) {
        uint64_t q[3];
        for (int c = 0; c < 3; ++c) q[c] = (uint64_t)fmin(fmax((x[3 * i + c] - x0[c]) * inv, 0.0), 2097151.0);
        kv[i].key = (spread3(q[0]) << 2) | (spread3(q[1]) << 1) | spread3(q[2]);
        kv[i].idx = (int)i;
    }
    qsort(kv, n, sizeof(fo_kv), cmp_kv);
    double *sx = malloc(sizeof(double) * 3 * n), *sg = malloc(sizeof(double) * 3 * n), *ss = malloc(sizeof(double) * n);
    for (int64_t i = 0; i < n; ++i) {
        memcpy(sx + 3 * i, x + 3 * kv[i].idx, 24);
        memcpy(sg + 3 * i, g + 3 * kv[i].idx, 24);
        ss[i] = sig[kv[i].idx];
    }
    fo_tree t;
    memset(&t, 0, sizeof(t));
    t.p = p; t.ncrit = ncrit; t.n = (int)n; t.theta = theta; t.x = sx; t.g = sg; t.s = ss; t.kernel = kernel;
    t.cap = (int)(8 * n / ncrit + 4096);
    t.cells = malloc(sizeof(fo_cell) * t.cap);
    fo_cell root = {0, (int)n, -1, -1, 0, 0, {ctr[0], ctr[1], ctr[2]}, 0.5 * side};
    t.cells[0] = root;
    t.ncells = 1;
    /* breadth first, children contiguous and in octant order (the order of fmm_split_emit_kernel) */
    for (int c = 0; c < t.ncells; ++c) {
        fo_cell cell = t.cells[c];
        int split = cell.count > ncrit;
        if (!split && leaf_sigmas > 0) {
            double smax = 0;
            for (int q = 0; q < cell.count; ++q) smax = fmax(smax, ss[cell.start + q]);
            split = cell.R > leaf_sigmas * smax && smax * 4096.0 > cell.R;
        }
        if (!split || cell.level >= FO_MAXLEVEL) continue;
        const int shift = 3 * (FO_MAXLEVEL - cell.level - 1);
        const uint64_t prefix = (kv[cell.start].key >> (shift + 3)) << 3;
        int lo_i = cell.start, k = 0;
        const double h = 0.5 * cell.R;
        for (int o = 0; o < 8; ++o) {
            int hi_i = o == 7 ? cell.start + cell.count : lower_bound(kv, lo_i, cell.start + cell.count, (prefix | (uint64_t)(o + 1)) << shift);
            if (hi_i > lo_i) {
                if (t.ncells == t.cap) {
                    t.cap *= 2;
                    t.cells = realloc(t.cells, sizeof(fo_cell) * t.cap);
                }
                fo_cell ch = {lo_i, hi_i - lo_i, c, -1, 0, cell.level + 1,
                              {cell.c[0] + ((o & 4) ? h : -h), cell.c[1] + ((o & 2) ? h : -h), cell.c[2] + ((o & 1) ? h : -h)}, h};
                if (k == 0) t.cells[c].child0 = t.ncells;
                t.cells[t.ncells++] = ch;
                ++k;
            }
            lo_i = hi_i;
        }
        t.cells[c].nchild = k;
    }
    t.M = calloc((size_t)t.ncells * 3 * FO_NCOEF, sizeof(cplx));
    t.L = calloc((size_t)t.ncells * 3 * FO_NCOEF, sizeof(cplx));
    t.U = calloc((size_t)3 * n, sizeof(double));
    t.J = calloc((size_t)9 * n, sizeof(double));
    const int P1 = p - 1;
    int nleaves = 0;
    /* upward: P2M at the leaves (charge q = Gamma / 4 pi), M2M to the parents (children have larger indices) */
    for (int c = t.ncells - 1; c >= 0; --c) {
        const fo_cell *cell = &t.cells[c];
        if (cell->nchild == 0) {
            ++nleaves;
            cplx Rh[FO_NCOEF];
            for (int q = 0; q < cell->count; ++q) {
                const double y[3] = {sx[3 * (cell->start + q)] - cell->c[0], sx[3 * (cell->start + q) + 1] - cell->c[1],
                                     sx[3 * (cell->start + q) + 2] - cell->c[2]};
                harmonics(y, P1, 0, Rh);
                for (int comp = 0; comp < 3; ++comp) {
                    const double ch = CONST4 * sg[3 * (cell->start + q) + comp];
                    cplx *Mc = t.M + ((size_t)c * 3 + comp) * FO_NCOEF;
                    for (int k = 0; k < p * p; ++k) Mc[k] += ch * conj(Rh[k]);
                }
            }
        }
        if (cell->parent >= 0) {   /* M_parent[n,m] += sum_{k,l} conj(R_k^l(d)) M_child[n-k, m-l],  d = c_child - c_parent */
            const fo_cell *par = &t.cells[cell->parent];
            const double d[3] = {cell->c[0] - par->c[0], cell->c[1] - par->c[1], cell->c[2] - par->c[2]};
            cplx Rh[FO_NCOEF];
            harmonics(d, P1, 0, Rh);
            for (int comp = 0; comp < 3; ++comp) {
                const cplx *Mc = t.M + ((size_t)c * 3 + comp) * FO_NCOEF;
                cplx *Mp = t.M + ((size_t)cell->parent * 3 + comp) * FO_NCOEF;
                for (int nn = 0; nn < p; ++nn)
                    for (int m = -nn; m <= nn; ++m) {
                        cplx acc = 0;
                        for (int k = 0; k <= nn; ++k)
                            for (int l = -k; l <= k; ++l)
                                if (abs(m - l) <= nn - k) acc += conj(Rh[idx_nm(k, l)]) * Mc[idx_nm(nn - k, m - l)];
                        Mp[idx_nm(nn, m)] += acc;
                    }
            }
        }
    }
    traverse(&t, 0, 0);
    /* downward: L2L to the children (increasing index), L2P at the leaves */
    for (int c = 0; c < t.ncells; ++c) {
        const fo_cell *cell = &t.cells[c];
        if (cell->parent >= 0) {   /* L_child[j,i] += sum_{k>=j,l} L_parent[k,l] conj(R_{k-j}^{l-i}(e)),  e = c_child - c_parent */
            const fo_cell *par = &t.cells[cell->parent];
            const double e[3] = {cell->c[0] - par->c[0], cell->c[1] - par->c[1], cell->c[2] - par->c[2]};
            cplx Rh[FO_NCOEF];
            harmonics(e, P1, 0, Rh);
            for (int comp = 0; comp < 3; ++comp) {
                const cplx *Lp = t.L + ((size_t)cell->parent * 3 + comp) * FO_NCOEF;
                cplx *Lc = t.L + ((size_t)c * 3 + comp) * FO_NCOEF;
                for (int jj = 0; jj < p; ++jj)
                    for (int i = -jj; i <= jj; ++i) {
                        cplx acc = 0;
                        for (int k = jj; k < p; ++k)
                            for (int l = -k; l <= k; ++l)
                                if (abs(l - i) <= k - jj) acc += Lp[idx_nm(k, l)] * conj(Rh[idx_nm(k - jj, l - i)]);
                        Lc[idx_nm(jj, i)] += acc;
                    }
            }
        }
        if (cell->nchild != 0) continue;
        for (int q = 0; q < cell->count; ++q) {
            const int ip = cell->start + q;
            const double y[3] = {sx[3 * ip] - cell->c[0], sx[3 * ip + 1] - cell->c[1], sx[3 * ip + 2] - cell->c[2]};
            cplx Rh[FO_NCOEF];
            harmonics(y, P1, 0, Rh);
            double G[3][3], H[3][3][3];   /* G[comp][a] = d_a psi_comp,  H[comp][a][b] = d_a d_b psi_comp */
            for (int comp = 0; comp < 3; ++comp) {
                const cplx *Lc = t.L + ((size_t)c * 3 + comp) * FO_NCOEF;
                cplx dz = 0, dp = 0, dm = 0, dzz = 0, dzp = 0, dzm = 0, dpp = 0, dmm = 0, dpm = 0;
                /* psi = sum conj(R_k^l) L_k^l; with D+ = dx + i dy, D- = dx - i dy acting on R:
                   Dz R_k^l = R_{k-1}^l, D+ R_k^l = R_{k-1}^{l+1}, D- R_k^l = -R_{k-1}^{l-1} (zero outside |l| <= k) */
#define RR(k, l) (((k) >= 0 && abs(l) <= (k)) ? Rh[idx_nm((k), (l))] : 0)
                for (int k = 1; k < p; ++k)
                    for (int l = -k; l <= k; ++l) {
                        const cplx Lkl = Lc[idx_nm(k, l)];
                        dz += conj(RR(k - 1, l)) * Lkl;
                        dp += conj(RR(k - 1, l + 1)) * Lkl;
                        dm += conj(-RR(k - 1, l - 1)) * Lkl;
                        if (k >= 2) {
                            dzz += conj(RR(k - 2, l)) * Lkl;
                            dzp += conj(RR(k - 2, l + 1)) * Lkl;
                            dzm += conj(-RR(k - 2, l - 1)) * Lkl;
                            dpp += conj(RR(k - 2, l + 2)) * Lkl;
                            dmm += conj(RR(k - 2, l - 2)) * Lkl;       /* (-1)(-1) */
                            dpm += conj(-RR(k - 2, l)) * Lkl;          /* D+ D- R_k^l = -R_{k-2}^{l} */
                        }
                    }
#undef RR
                /* conj(D+ R) = (dx - i dy) conj(R): with f = sum conj(R) L real, dx f = Re-part combos:
                   A = sum conj(D+ R) L = (dx - i dy) f,  B = sum conj(D- R) L = (dx + i dy) f  */
                const cplx A = dp, B = dm;
                G[comp][0] = creal(0.5 * (A + B));
                G[comp][1] = creal((B - A) / (2.0 * I));
                G[comp][2] = creal(dz);
                /* second derivatives: App = (dx - i dy)^2 f, Bmm = (dx + i dy)^2 f, Cpm = (dx - i dy)(dx + i dy) f = (dxx + dyy) f */
                const cplx App = dpp, Bmm = dmm, Cpm = dpm;
                const double fxx_m_fyy = creal(0.5 * (App + Bmm));        /* dxx - dyy */
                const double fxy = creal((Bmm - App) / (4.0 * I));        /* dxy */
                const double lap2 = creal(Cpm);                            /* dxx + dyy */
                H[comp][0][0] = 0.5 * (lap2 + fxx_m_fyy);
                H[comp][1][1] = 0.5 * (lap2 - fxx_m_fyy);
                H[comp][0][1] = H[comp][1][0] = fxy;
                H[comp][2][2] = creal(dzz);
                const cplx Az = dzp, Bz = dzm;                             /* (dx - i dy) dz f, (dx + i dy) dz f */
                H[comp][0][2] = H[comp][2][0] = creal(0.5 * (Az + Bz));
                H[comp][1][2] = H[comp][2][1] = creal((Bz - Az) / (2.0 * I));
            }
            /* U = curl psi;  J[i + 3 j] = d_j U_i */
            t.U[3 * ip + 0] += G[2][1] - G[1][2];
            t.U[3 * ip + 1] += G[0][2] - G[2][0];
            t.U[3 * ip + 2] += G[1][0] - G[0][1];
            for (int jd = 0; jd < 3; ++jd) {
                t.J[9 * ip + 0 + 3 * jd] += H[2][jd][1] - H[1][jd][2];
                t.J[9 * ip + 1 + 3 * jd] += H[0][jd][2] - H[2][jd][0];
                t.J[9 * ip + 2 + 3 * jd] += H[1][jd][0] - H[0][jd][1];
            }
        }
    }
    for (int64_t i = 0; i < n; ++i) {
        memcpy(U + 3 * kv[i].idx, t.U + 3 * i, 24);
        memcpy(J + 9 * kv[i].idx, t.J + 9 * i, 72);
    }
    if (stats) {
        stats[0] = t.ncells; stats[1] = nleaves; stats[2] = t.n_m2l; stats[3] = t.n_p2p;
    }
    free(kv); free(sx); free(sg); free(ss); free(t.cells); free(t.M); free(t.L); free(t.U); free(t.J);
    return 0;
}
