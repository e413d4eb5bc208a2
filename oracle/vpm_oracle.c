/*
 * vpm_oracle.c — CPU restatement of the rVPM hot path.  TEST INFRASTRUCTURE ONLY; see vpm_oracle.h
 * for scope, provenance and the "parity unpinned" statement.
 *
 * Every function cites what it follows:
 *   REF  = file:line in /root/reference (read directly),
 *   DOC  = the reference's docs/src/theory/rvpm.md,
 *   A.n  = SURVEY.md Appendix A section n (published FLOWVPM algorithm; UPSTREAM-RECALL).
 * Expressions are kept in the reference's own operation order (no algebraic re-association), so this
 * file is also the "reference form" against which the 86 flop/interaction figure is counted.
 */
#include "vpm_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* REF src/FLOWUnsteady_processing_force.jl:903 — the literal the reference uses for 1/(4 pi). */
static const double CONST4 = 0.07957747154594767;
/* A.3: 1/(2 pi)^(3/2), sqrt(2/pi), sqrt(2) */
static const double CONST1 = 0.06349363593424097;
static const double CONST2 = 0.7978845608028654;
static const double SQR2 = 1.4142135623730951;

int32_t vpmo_num_threads(void) {
#ifdef _OPENMP
    return (int32_t)omp_get_max_threads();
#else
    return 1;
#endif
}

void vpmo_set_num_threads(int32_t n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

void vpmo_default_schemes(vpmo_schemes *s) {
    /* REF src/FLOWUnsteady_simulation.jl:36-44 */
    memset(s, 0, sizeof(*s));
    s->kernel = VPMO_KERNEL_GAUSSIANERF;
    s->f = 0.0;
    s->g = 1.0 / 5.0;
    s->transposed = 1;
    s->relaxation = VPMO_RELAX_PEDRIZZETTI;
    s->rlxf = 0.3;
    s->sfs = VPMO_SFS_NONE;
    s->alpha = 0.999;
    s->sfs_rlxf = 0.005;
    s->minC = 0.0;
    s->maxC = 1.0;
    s->Cs = 1.0;
    s->force_positive = 0;
    s->clippings = 0;
    s->controls = 0;
    s->viscous = VPMO_VISCOUS_INVISCID;
    s->nu = 0.0;
    s->integration = VPMO_INTEGRATION_RK3;
    s->cs_sgm0 = 0.0;
    s->cs_beta = 1.5;
    s->cs_itmax = 15;
    s->cs_tol = 1e-3;
}

/* ------------------------------------------------------------------ kernels (A.3) ------------ */

void vpmo_g_dgdr(int32_t kernel, double r, double *g, double *dg) {
    switch (kernel) {
    case VPMO_KERNEL_GAUSSIANERF: {
        /* g = erf(r/sqrt2) - sqrt(2/pi) r exp(-r^2/2);  g' = sqrt(2/pi) r^2 exp(-r^2/2) */
        double aux = CONST2 * r * exp(-r * r / 2);
        *g = erf(r / SQR2) - aux;
        *dg = r * aux;
        break;
    }
    case VPMO_KERNEL_WINCKELMANS: {
        /* g = r^3 (r^2 + 2.5)/(r^2+1)^2.5;  g' = 7.5 r^2/(r^2+1)^3.5 */
        double aux0 = pow(r * r + 1, 2.5);
        *g = r * r * r * (r * r + 2.5) / aux0;
        *dg = 7.5 * r * r / (aux0 * (r * r + 1));
        break;
    }
    case VPMO_KERNEL_GAUSSIAN: {
        /* g = 1 - exp(-r^3);  g' = 3 r^2 exp(-r^3) */
        double aux = exp(-(r * r * r));
        *g = 1 - aux;
        *dg = 3 * r * r * aux;
        break;
    }
    default: /* singular */
        *g = 1.0;
        *dg = 0.0;
        break;
    }
}

double vpmo_zeta(int32_t kernel, double r) {
    switch (kernel) {
    case VPMO_KERNEL_GAUSSIANERF:
        return CONST1 * exp(-r * r / 2);
    case VPMO_KERNEL_WINCKELMANS:
        return CONST4 * 7.5 / pow(r * r + 1, 3.5);
    case VPMO_KERNEL_GAUSSIAN:
        return 3 * CONST4 * exp(-(r * r * r));
    default: /* singular: Dirac delta, represented as in upstream by 1 at r == 0 */
        return r == 0 ? 1.0 : 0.0;
    }
}

/* ------------------------------------------------------------------ UJ_direct (A.2) ---------- */

#define UJ_PAIR_BODY(ACC_T)                                                                       \
    double dX1 = xt[3 * i + 0] - xs[3 * j + 0];                                                   \
    double dX2 = xt[3 * i + 1] - xs[3 * j + 1];                                                   \
    double dX3 = xt[3 * i + 2] - xs[3 * j + 2];                                                   \
    double r = sqrt(dX1 * dX1 + dX2 * dX2 + dX3 * dX3);                                           \
    if (r != 0) { /* REF processing_force.jl:895 */                                               \
        double g_sgm, dg_sgmdr;                                                                   \
        vpmo_g_dgdr(kernel, r / sig[j], &g_sgm, &dg_sgmdr); /* sigma of the SOURCE, REF :898 */   \
        double G1 = gs[3 * j + 0], G2 = gs[3 * j + 1], G3 = gs[3 * j + 2];                        \
        /* K x Gamma_p, REF :903-905 */                                                           \
        double r3 = r * r * r;                                                                    \
        double crss1 = -CONST4 / r3 * (dX2 * G3 - dX3 * G2);                                      \
        double crss2 = -CONST4 / r3 * (dX3 * G1 - dX1 * G3);                                      \
        double crss3 = -CONST4 / r3 * (dX1 * G2 - dX2 * G1);                                      \
        /* U = sum g_sigma K x Gamma */                                                           \
        u[0] += (ACC_T)(g_sgm * crss1);                                                           \
        u[1] += (ACC_T)(g_sgm * crss2);                                                           \
        u[2] += (ACC_T)(g_sgm * crss3);                                                           \
        /* du/dx_j = (dx_j g'/(sigma r) - 3 dx_j g/r^2) K x Gamma */                              \
        double aux = dg_sgmdr / (sig[j] * r) - 3 * g_sgm / (r * r);                               \
        jac[0] += (ACC_T)(aux * crss1 * dX1);                                                     \
        jac[1] += (ACC_T)(aux * crss2 * dX1);                                                     \
        jac[2] += (ACC_T)(aux * crss3 * dX1);                                                     \
        jac[3] += (ACC_T)(aux * crss1 * dX2);                                                     \
        jac[4] += (ACC_T)(aux * crss2 * dX2);                                                     \
        jac[5] += (ACC_T)(aux * crss3 * dX2);                                                     \
        jac[6] += (ACC_T)(aux * crss1 * dX3);                                                     \
        jac[7] += (ACC_T)(aux * crss2 * dX3);                                                     \
        jac[8] += (ACC_T)(aux * crss3 * dX3);                                                     \
        /* Kronecker-delta term: -g/(4 pi r^3) delta_ij x Gamma */                                \
        aux = -CONST4 * g_sgm / r3;                                                               \
        jac[1] -= (ACC_T)(aux * G3); /* J[2,1] */                                                 \
        jac[2] += (ACC_T)(aux * G2); /* J[3,1] */                                                 \
        jac[3] += (ACC_T)(aux * G3); /* J[1,2] */                                                 \
        jac[5] -= (ACC_T)(aux * G1); /* J[3,2] */                                                 \
        jac[6] -= (ACC_T)(aux * G2); /* J[1,3] */                                                 \
        jac[7] += (ACC_T)(aux * G1); /* J[2,3] */                                                 \
    }

void vpmo_uj_direct(int32_t kernel, int64_t ns, const double *xs, const double *gs, const double *sig,
                    int64_t nt, const double *xt, double *U, double *J, int32_t accum) {
    if (accum == 0) {
#pragma omp parallel for schedule(static) if (nt >= 256)   /* (the FMM oracle calls this per leaf pair) */
        for (int64_t i = 0; i < nt; ++i) {
            double u[3] = {U[3 * i], U[3 * i + 1], U[3 * i + 2]};
            double jac[9];
            for (int k = 0; k < 9; ++k) jac[k] = J[9 * i + k];
            for (int64_t j = 0; j < ns; ++j) {
                UJ_PAIR_BODY(double)
            }
            for (int k = 0; k < 3; ++k) U[3 * i + k] = u[k];
            for (int k = 0; k < 9; ++k) J[9 * i + k] = jac[k];
        }
    } else {
#pragma omp parallel for schedule(static) if (nt >= 256)   /* (the FMM oracle calls this per leaf pair) */
        for (int64_t i = 0; i < nt; ++i) {
            long double u[3] = {U[3 * i], U[3 * i + 1], U[3 * i + 2]};
            long double jac[9];
            for (int k = 0; k < 9; ++k) jac[k] = J[9 * i + k];
            for (int64_t j = 0; j < ns; ++j) {
                UJ_PAIR_BODY(long double)
            }
            for (int k = 0; k < 3; ++k) U[3 * i + k] = (double)u[k];
            for (int k = 0; k < 9; ++k) J[9 * i + k] = (double)jac[k];
        }
    }
}

/* ------------------------------------------------------------------ Estr_direct (A.4) -------- */

#define ESTR_PAIR_BODY(ACC_T)                                                                     \
    double G1 = gs[3 * j + 0], G2 = gs[3 * j + 1], G3 = gs[3 * j + 2];                            \
    const double *Jp = Jt + 9 * i;                                                                \
    const double *Jq = Js + 9 * j;                                                                \
    double S1, S2, S3;                                                                            \
    if (transposed) { /* S_k = sum_l (Jp[l,k] - Jq[l,k]) Gamma_q,l */                             \
        S1 = (Jp[0] - Jq[0]) * G1 + (Jp[1] - Jq[1]) * G2 + (Jp[2] - Jq[2]) * G3;                  \
        S2 = (Jp[3] - Jq[3]) * G1 + (Jp[4] - Jq[4]) * G2 + (Jp[5] - Jq[5]) * G3;                  \
        S3 = (Jp[6] - Jq[6]) * G1 + (Jp[7] - Jq[7]) * G2 + (Jp[8] - Jq[8]) * G3;                  \
    } else { /* S_k = sum_l (Jp[k,l] - Jq[k,l]) Gamma_q,l */                                      \
        S1 = (Jp[0] - Jq[0]) * G1 + (Jp[3] - Jq[3]) * G2 + (Jp[6] - Jq[6]) * G3;                  \
        S2 = (Jp[1] - Jq[1]) * G1 + (Jp[4] - Jq[4]) * G2 + (Jp[7] - Jq[7]) * G3;                  \
        S3 = (Jp[2] - Jq[2]) * G1 + (Jp[5] - Jq[5]) * G2 + (Jp[8] - Jq[8]) * G3;                  \
    }                                                                                             \
    double dX1 = xt[3 * i + 0] - xs[3 * j + 0];                                                   \
    double dX2 = xt[3 * i + 1] - xs[3 * j + 1];                                                   \
    double dX3 = xt[3 * i + 2] - xs[3 * j + 2];                                                   \
    double r = sqrt(dX1 * dX1 + dX2 * dX2 + dX3 * dX3);                                           \
    double zeta_sgm = 1 / (sig[j] * sig[j] * sig[j]) * vpmo_zeta(kernel, r / sig[j]);             \
    e[0] += (ACC_T)(zeta_sgm * S1);                                                               \
    e[1] += (ACC_T)(zeta_sgm * S2);                                                               \
    e[2] += (ACC_T)(zeta_sgm * S3);

void vpmo_estr_direct(int32_t kernel, int32_t transposed, int64_t ns, const double *xs, const double *gs,
                      const double *sig, const double *Js, int64_t nt, const double *xt, const double *Jt,
                      double *SFS, int32_t accum) {
    if (accum == 0) {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nt; ++i) {
            double e[3] = {SFS[3 * i], SFS[3 * i + 1], SFS[3 * i + 2]};
            for (int64_t j = 0; j < ns; ++j) {
                ESTR_PAIR_BODY(double)
            }
            for (int k = 0; k < 3; ++k) SFS[3 * i + k] = e[k];
        }
    } else {
#pragma omp parallel for schedule(static)
        for (int64_t i = 0; i < nt; ++i) {
            long double e[3] = {SFS[3 * i], SFS[3 * i + 1], SFS[3 * i + 2]};
            for (int64_t j = 0; j < ns; ++j) {
                ESTR_PAIR_BODY(long double)
            }
            for (int k = 0; k < 3; ++k) SFS[3 * i + k] = (double)e[k];
        }
    }
}

/* ------------------------------------------------------------------ zeta pass / RBF (A.8) ---- */

void vpmo_zeta_direct(int32_t kernel, int64_t ns, const double *xs, const double *vs, const double *sig, int64_t nt,
                      const double *xt, double *out) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < nt; ++i) {
        long double a[3] = {out[3 * i], out[3 * i + 1], out[3 * i + 2]};
        for (int64_t j = 0; j < ns; ++j) {
            double dX1 = xt[3 * i + 0] - xs[3 * j + 0];
            double dX2 = xt[3 * i + 1] - xs[3 * j + 1];
            double dX3 = xt[3 * i + 2] - xs[3 * j + 2];
            double r = sqrt(dX1 * dX1 + dX2 * dX2 + dX3 * dX3);
            double z = 1 / (sig[j] * sig[j] * sig[j]) * vpmo_zeta(kernel, r / sig[j]);
            a[0] += z * vs[3 * j + 0];
            a[1] += z * vs[3 * j + 1];
            a[2] += z * vs[3 * j + 2];
        }
        for (int k = 0; k < 3; ++k) out[3 * i + k] = (double)a[k];
    }
}

/* ------------------------------------------------------------------ _Ffv_direct (REF) -------- */

void vpmo_ffv_direct(int32_t kernel, int64_t nb, const double *xb, const double *gb, const double *sb,
                     int64_t nf, const double *xf, const double *gf, double *M6) {
    /* REF src/FLOWUnsteady_processing_force.jl:879-929.  Parallel over bound vortices (:881). */
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < nb; ++b) {
        double M[6] = {0, 0, 0, 0, 0, 0}; /* :884 */
        double B1 = gb[3 * b], B2 = gb[3 * b + 1], B3 = gb[3 * b + 2];
        for (int64_t f = 0; f < nf; ++f) {
            double dX1 = xf[3 * f + 0] - xb[3 * b + 0]; /* :890-892, dx = xf - xb */
            double dX2 = xf[3 * f + 1] - xb[3 * b + 1];
            double dX3 = xf[3 * f + 2] - xb[3 * b + 2];
            double r = sqrt(dX1 * dX1 + dX2 * dX2 + dX3 * dX3);
            if (r != 0) {
                double g_sgm, dg_sgmdr;
                vpmo_g_dgdr(kernel, r / sb[b], &g_sgm, &dg_sgmdr); /* :898 */
                double r3 = r * r * r;
                double F1 = gf[3 * f], F2 = gf[3 * f + 1], F3 = gf[3 * f + 2];
                /* :903-905 */
                double U1 = -g_sgm * CONST4 / r3 * (dX2 * B3 - dX3 * B2);
                double U2 = -g_sgm * CONST4 / r3 * (dX3 * B1 - dX1 * B3);
                double U3 = -g_sgm * CONST4 / r3 * (dX1 * B2 - dX2 * B1);
                /* :908-910 */
                M[0] += U2 * F3 - U3 * F2;
                M[1] += U3 * F1 - U1 * F3;
                M[2] += U1 * F2 - U2 * F1;
                /* :915-917 */
                double crss1 = B2 * F3 - B3 * F2;
                double crss2 = B3 * F1 - B1 * F3;
                double crss3 = B1 * F2 - B2 * F1;
                /* :920-922 */
                M[3] += g_sgm * CONST4 / r3 * (dX2 * crss3 - dX3 * crss2);
                M[4] += g_sgm * CONST4 / r3 * (dX3 * crss1 - dX1 * crss3);
                M[5] += g_sgm * CONST4 / r3 * (dX1 * crss2 - dX2 * crss1);
            }
        }
        for (int k = 0; k < 6; ++k) M6[6 * b + k] = M[k];
    }
}

/* ------------------------------------------------------------------ field-level helpers ------ */

#define PCOL(P, i) ((P) + (int64_t)VPMO_NFIELDS * (i))

static int is_static(const double *p) { return p[VPMO_STATIC] > 0; } /* REF simulation.jl:509 */

void vpmo_reset_particles(double *P, int64_t np) {
    /* A.2: vpm._reset_particles (REF use: processing_force.jl:237): U, J, PSE <- 0 */
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        for (int k = 0; k < 3; ++k) p[VPMO_U + k] = 0;
        for (int k = 0; k < 9; ++k) p[VPMO_J + k] = 0;
        for (int k = 0; k < 3; ++k) p[VPMO_PSE + k] = 0;
    }
}

void vpmo_reset_particles_sfs(double *P, int64_t np) {
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        for (int k = 0; k < 3; ++k) p[VPMO_SFS + k] = 0;
    }
}

void vpmo_field_uj(double *P, int64_t np, const vpmo_schemes *s, int32_t reset, int32_t reset_sfs,
                   int32_t sfs) {
    if (reset) vpmo_reset_particles(P, np);
    if (reset_sfs) vpmo_reset_particles_sfs(P, np);
    if (np <= 0) return;
    double *x = (double *)malloc(sizeof(double) * 3 * np);
    double *gm = (double *)malloc(sizeof(double) * 3 * np);
    double *sg = (double *)malloc(sizeof(double) * np);
    double *U = (double *)malloc(sizeof(double) * 3 * np);
    double *J = (double *)malloc(sizeof(double) * 9 * np);
    for (int64_t i = 0; i < np; ++i) {
        const double *p = PCOL(P, i);
        for (int k = 0; k < 3; ++k) x[3 * i + k] = p[VPMO_X + k];
        for (int k = 0; k < 3; ++k) gm[3 * i + k] = p[VPMO_GAMMA + k];
        sg[i] = p[VPMO_SIGMA];
        for (int k = 0; k < 3; ++k) U[3 * i + k] = p[VPMO_U + k];
        for (int k = 0; k < 9; ++k) J[9 * i + k] = p[VPMO_J + k];
    }
    /* sources and targets: every particle, statics included (they induce and are probed) */
    vpmo_uj_direct(s->kernel, np, x, gm, sg, np, x, U, J, 0);
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        for (int k = 0; k < 3; ++k) p[VPMO_U + k] = U[3 * i + k];
        for (int k = 0; k < 9; ++k) p[VPMO_J + k] = J[9 * i + k];
    }
    if (sfs) {
        double *E = (double *)malloc(sizeof(double) * 3 * np);
        for (int64_t i = 0; i < np; ++i)
            for (int k = 0; k < 3; ++k) E[3 * i + k] = PCOL(P, i)[VPMO_SFS + k];
        vpmo_estr_direct(s->kernel, s->transposed, np, x, gm, sg, J, np, x, J, E, 0);
        for (int64_t i = 0; i < np; ++i)
            for (int k = 0; k < 3; ++k) PCOL(P, i)[VPMO_SFS + k] = E[3 * i + k];
        free(E);
    }
    free(x);
    free(gm);
    free(sg);
    free(U);
    free(J);
}

/* S = (Gamma . grad') u (transposed) or (Gamma . grad) u  — A.6 */
static void stretching(const double *p, int transposed, double S[3]) {
    const double *J = p + VPMO_J;
    const double *G = p + VPMO_GAMMA;
    if (transposed) {
        S[0] = J[0] * G[0] + J[1] * G[1] + J[2] * G[2];
        S[1] = J[3] * G[0] + J[4] * G[1] + J[5] * G[2];
        S[2] = J[6] * G[0] + J[7] * G[1] + J[8] * G[2];
    } else {
        S[0] = J[0] * G[0] + J[3] * G[1] + J[6] * G[2];
        S[1] = J[1] * G[0] + J[4] * G[1] + J[7] * G[2];
        S[2] = J[2] * G[0] + J[5] * G[1] + J[8] * G[2];
    }
}

static double sgn(double x) { return (x > 0) - (x < 0); }

/* A.5 clipping_backscatter (DOC rvpm.md:296): clip when C_d Gamma . E_str < 0 */
static int clipping_backscatter(const double *p) {
    const double *G = p + VPMO_GAMMA, *E = p + VPMO_SFS;
    return p[VPMO_C] * (G[0] * E[0] + G[1] * E[1] + G[2] * E[2]) < 0;
}

/* A.5 control_directional: project the SFS term onto Gamma */
static void control_directional(double *p) {
    double *G = p + VPMO_GAMMA, *E = p + VPMO_SFS;
    double aux = E[0] * G[0] + E[1] * G[1] + E[2] * G[2];
    aux /= (G[0] * G[0] + G[1] * G[1] + G[2] * G[2]);
    for (int k = 0; k < 3; ++k) E[k] += -E[k] + aux * G[k];
}

/* A.5 control_magnitude: limit the SFS term so it cannot reverse Gamma within one step */
static void control_magnitude(double *p, const vpmo_schemes *s, double t, int64_t nt, double zeta0) {
    if (nt == 0) return;
    if (p[VPMO_C] == 0) return;
    double *G = p + VPMO_GAMMA, *E = p + VPMO_SFS;
    double deltat = t / (double)nt;
    double sg = p[VPMO_SIGMA];
    double aux = E[0] * G[0] + E[1] * G[1] + E[2] * G[2];
    aux /= (G[0] * G[0] + G[1] * G[1] + G[2] * G[2]);
    aux -= (1 + 3 * s->f) * (zeta0 / (sg * sg * sg)) / deltat / p[VPMO_C];
    if (aux > 0)
        for (int k = 0; k < 3; ++k) E[k] += -aux * G[k];
}

static void apply_clippings_controls(double *P, int64_t np, const vpmo_schemes *s, double t, int64_t nt,
                                     double zeta0) {
    if (s->clippings & VPMO_CLIP_BACKSCATTER)
        for (int64_t i = 0; i < np; ++i) {
            double *p = PCOL(P, i);
            if (!is_static(p) && clipping_backscatter(p)) p[VPMO_C] *= 0;
        }
    if (s->controls & VPMO_CTRL_DIRECTIONAL)
        for (int64_t i = 0; i < np; ++i) {
            double *p = PCOL(P, i);
            if (!is_static(p)) control_directional(p);
        }
    if (s->controls & VPMO_CTRL_MAGNITUDE)
        for (int64_t i = 0; i < np; ++i) {
            double *p = PCOL(P, i);
            if (!is_static(p)) control_magnitude(p, s, t, nt, zeta0);
        }
}

/* A.5 dynamicprocedure_pseudo3level (DOC rvpm.md:264-296) */
static void dynamic_pseudo3level(double *P, int64_t np, const vpmo_schemes *s, double zeta0) {
    const double alpha = s->alpha, rlxf = s->sfs_rlxf, minC = s->minC, maxC = s->maxC;
    /* test filter: sigma <- alpha sigma for the non-static particles */
    for (int64_t i = 0; i < np; ++i)
        if (!is_static(PCOL(P, i))) PCOL(P, i)[VPMO_SIGMA] *= alpha;
    vpmo_field_uj(P, np, s, 1, 1, 1);
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        if (is_static(p)) continue;
        for (int k = 0; k < 9; ++k) p[VPMO_M + k] = 0;
        double S[3];
        stretching(p, s->transposed, S);
        for (int k = 0; k < 3; ++k) p[VPMO_M + k] = S[k];                   /* M[:,1] */
        for (int k = 0; k < 3; ++k) p[VPMO_M + 3 + k] = p[VPMO_SFS + k];     /* M[:,2] */
    }
    /* domain filter: restore sigma */
    for (int64_t i = 0; i < np; ++i)
        if (!is_static(PCOL(P, i))) PCOL(P, i)[VPMO_SIGMA] /= alpha;
    vpmo_field_uj(P, np, s, 1, 1, 1);
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        if (is_static(p)) continue;
        double S[3];
        stretching(p, s->transposed, S);
        for (int k = 0; k < 3; ++k) p[VPMO_M + k] -= S[k];
        for (int k = 0; k < 3; ++k) p[VPMO_M + 3 + k] -= p[VPMO_SFS + k];
    }
    /* coefficient */
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        if (is_static(p)) continue;
        double *G = p + VPMO_GAMMA, *M = p + VPMO_M, *C = p + VPMO_C;
        double sg = p[VPMO_SIGMA];
        double nume = M[0] * G[0] + M[1] * G[1] + M[2] * G[2];
        nume *= 3 * alpha - 2;
        double deno = M[3] * G[0] + M[4] * G[1] + M[5] * G[2];
        deno /= zeta0 / (sg * sg * sg);
        if (C[2] == 0) {
            C[2] = deno;
            if (C[2] == 0) C[2] = DBL_EPSILON;
        }
        nume = rlxf * nume + (1 - rlxf) * C[1];
        deno = rlxf * deno + (1 - rlxf) * C[2];
        if (fabs(nume / deno) > maxC) {
            if (fabs(deno) < fabs(C[2])) deno = sgn(deno) * fabs(C[2]);
            nume = sgn(nume) * fabs(deno) * maxC;
        } else if (fabs(nume / deno) < minC) {
            nume = sgn(nume) * fabs(deno) * minC;
        }
        C[1] = nume;
        C[2] = deno;
        C[0] = C[1] / C[2];
        if (s->force_positive) C[0] = fabs(C[0]);
        for (int k = 0; k < 9; ++k) M[k] = 0;
    }
}

void vpmo_field_sfs(double *P, int64_t np, const vpmo_schemes *s, double a, double b, double t, int64_t nt) {
    (void)b;
    const double zeta0 = vpmo_zeta(s->kernel, 0.0);
    const int first = (a == 1 || a == 0); /* Euler step or first RK substep (A.5) */
    switch (s->sfs) {
    case VPMO_SFS_NONE:
        vpmo_field_uj(P, np, s, 1, 0, 0);
        break;
    case VPMO_SFS_CONSTANT:
        vpmo_field_uj(P, np, s, 1, 1, 1);
        if (first) {
            for (int64_t i = 0; i < np; ++i)
                if (!is_static(PCOL(P, i))) PCOL(P, i)[VPMO_C] = s->Cs;
            apply_clippings_controls(P, np, s, t, nt, zeta0);
        }
        break;
    default: /* dynamic */
        if (first) {
            dynamic_pseudo3level(P, np, s, zeta0);
            apply_clippings_controls(P, np, s, t, nt, zeta0);
        } else {
            vpmo_field_uj(P, np, s, 1, 1, 1);
        }
        break;
    }
}

/* A.7 relaxation (DOC rvpm.md:363) */
void vpmo_relax_particle(double *p, int32_t relaxation, double rlxf) {
    if (relaxation == VPMO_RELAX_NONE) return;
    const double *J = p + VPMO_J;
    double *G = p + VPMO_GAMMA;
    /* omega = curl u: (J[3,2]-J[2,3], J[1,3]-J[3,1], J[2,1]-J[1,2]) ; J[i,j] at i + 3 j (0-based) */
    double w1 = J[2 + 3 * 1] - J[1 + 3 * 2];
    double w2 = J[0 + 3 * 2] - J[2 + 3 * 0];
    double w3 = J[1 + 3 * 0] - J[0 + 3 * 1];
    double nrmw = sqrt(w1 * w1 + w2 * w2 + w3 * w3);
    double nrmGamma = sqrt(G[0] * G[0] + G[1] * G[1] + G[2] * G[2]);
    if (relaxation == VPMO_RELAX_PEDRIZZETTI) {
        G[0] = (1 - rlxf) * G[0] + rlxf * nrmGamma * w1 / nrmw;
        G[1] = (1 - rlxf) * G[1] + rlxf * nrmGamma * w2 / nrmw;
        G[2] = (1 - rlxf) * G[2] + rlxf * nrmGamma * w3 / nrmw;
    } else {
        double b2 = 1 - 2 * (1 - rlxf) * rlxf * (1 - (G[0] * w1 + G[1] * w2 + G[2] * w3) / (nrmGamma * nrmw));
        G[0] = (1 - rlxf) * G[0] + rlxf * nrmGamma * w1 / nrmw;
        G[1] = (1 - rlxf) * G[1] + rlxf * nrmGamma * w2 / nrmw;
        G[2] = (1 - rlxf) * G[2] + rlxf * nrmGamma * w3 / nrmw;
        double sb = sqrt(b2);
        G[0] /= sb;
        G[1] /= sb;
        G[2] /= sb;
    }
}

/* A.6 one low-storage substep for one non-static particle (DOC rvpm.md:107-235,364), followed by the
 * A.8 core-spreading sigma update.  q-storage: qU = M[:,1], qGamma = M[:,2], q_sigma2 = M[1,3],
 * q_sigma = M[2,3].  Euler is (a, b) = (0, 1). */
void vpmo_update_particle(double *p, const vpmo_schemes *s, double a, double b, double dt,
                          const double *Uinf, double zeta0) {
    double *X = p + VPMO_X, *G = p + VPMO_GAMMA, *M = p + VPMO_M;
    const double *U = p + VPMO_U, *E = p + VPMO_SFS;
    const double f = s->f, g = s->g;
    const double C = p[VPMO_C];
    for (int k = 0; k < 3; ++k) {
        M[k] = a * M[k] + dt * (U[k] + Uinf[k]);
        X[k] += b * M[k];
    }
    double S[3];
    stretching(p, s->transposed, S);
    double sg = p[VPMO_SIGMA];
    double sg3z = sg * sg * sg / zeta0;
    double Z = (f + g) / (1 + 3 * f) * (S[0] * G[0] + S[1] * G[1] + S[2] * G[2]);
    Z -= f / (1 + 3 * f) * (C * E[0] * G[0] + C * E[1] * G[1] + C * E[2] * G[2]) * sg3z;
    Z /= G[0] * G[0] + G[1] * G[1] + G[2] * G[2];
    for (int k = 0; k < 3; ++k) M[3 + k] = a * M[3 + k] + dt * (S[k] - 3 * Z * G[k] - C * E[k] * sg3z);
    M[7] = a * M[7] - dt * (sg * Z);
    for (int k = 0; k < 3; ++k) G[k] += b * M[3 + k];
    p[VPMO_SIGMA] += b * M[7];
    if (s->viscous == VPMO_VISCOUS_CORESPREADING) {
        M[6] = a * M[6] + dt * 2 * s->nu;
        p[VPMO_SIGMA] = sqrt(p[VPMO_SIGMA] * p[VPMO_SIGMA] + b * M[6]);
    }
}

int32_t vpmo_corespreading_reset(double *P, int64_t np, int32_t kernel, double sgm0, double beta, int32_t itmax, double tol,
                                 double *residual3) {
    int need = 0;
    for (int64_t i = 0; i < np; ++i) {
        const double *p = PCOL(P, i);
        if (!is_static(p) && p[VPMO_SIGMA] / sgm0 > beta) need = 1;
    }
    if (residual3) residual3[0] = residual3[1] = residual3[2] = 0;
    if (!need || np <= 0) return 0;
    double *x = (double *)malloc(sizeof(double) * 3 * np), *v = (double *)malloc(sizeof(double) * 3 * np);
    double *sg = (double *)malloc(sizeof(double) * np), *b = (double *)calloc(3 * np, sizeof(double));
    double *r = (double *)malloc(sizeof(double) * 3 * np), *d = (double *)malloc(sizeof(double) * 3 * np);
    double *Ad = (double *)malloc(sizeof(double) * 3 * np);
    for (int64_t i = 0; i < np; ++i) {
        const double *p = PCOL(P, i);
        for (int k = 0; k < 3; ++k) { x[3 * i + k] = p[VPMO_X + k]; v[3 * i + k] = p[VPMO_GAMMA + k]; }
        sg[i] = p[VPMO_SIGMA];
    }
    /* target vorticity with the spread cores */
    vpmo_zeta_direct(kernel, np, x, v, sg, np, x, b);
    for (int64_t i = 0; i < np; ++i) {
        double *p = PCOL(P, i);
        for (int k = 0; k < 3; ++k) p[VPMO_W + k] = b[3 * i + k];
        if (!is_static(p)) { p[VPMO_SIGMA] = sgm0; sg[i] = sgm0; }
    }
    /* CG on the non-static unknowns; statics contribute A_static Gamma_static to both sides (kept inside A v below and
     * never updated because their search direction is zero) */
    memset(Ad, 0, sizeof(double) * 3 * np);
    vpmo_zeta_direct(kernel, np, x, v, sg, np, x, Ad);
    double rr[3] = {0, 0, 0};
    for (int64_t i = 0; i < np; ++i) {
        int st = is_static(PCOL(P, i));
        for (int k = 0; k < 3; ++k) {
            r[3 * i + k] = st ? 0.0 : b[3 * i + k] - Ad[3 * i + k];
            d[3 * i + k] = r[3 * i + k];
            rr[k] += r[3 * i + k] * r[3 * i + k];
        }
    }
    int32_t it = 0;
    for (; it < itmax; ++it) {
        if (sqrt(rr[0]) < tol && sqrt(rr[1]) < tol && sqrt(rr[2]) < tol) break;
        memset(Ad, 0, sizeof(double) * 3 * np);
        vpmo_zeta_direct(kernel, np, x, d, sg, np, x, Ad);
        double dAd[3] = {0, 0, 0};
        for (int64_t i = 0; i < np; ++i)
            if (!is_static(PCOL(P, i)))
                for (int k = 0; k < 3; ++k) dAd[k] += d[3 * i + k] * Ad[3 * i + k];
        double alpha[3], rrn[3] = {0, 0, 0};
        for (int k = 0; k < 3; ++k) alpha[k] = dAd[k] != 0 ? rr[k] / dAd[k] : 0.0;
        for (int64_t i = 0; i < np; ++i) {
            if (is_static(PCOL(P, i))) continue;
            for (int k = 0; k < 3; ++k) {
                v[3 * i + k] += alpha[k] * d[3 * i + k];
                r[3 * i + k] -= alpha[k] * Ad[3 * i + k];
                rrn[k] += r[3 * i + k] * r[3 * i + k];
            }
        }
        for (int64_t i = 0; i < np; ++i) {
            if (is_static(PCOL(P, i))) continue;
            for (int k = 0; k < 3; ++k) d[3 * i + k] = r[3 * i + k] + (rr[k] != 0 ? rrn[k] / rr[k] : 0.0) * d[3 * i + k];
        }
        for (int k = 0; k < 3; ++k) rr[k] = rrn[k];
    }
    for (int64_t i = 0; i < np; ++i)
        for (int k = 0; k < 3; ++k) PCOL(P, i)[VPMO_GAMMA + k] = v[3 * i + k];
    if (residual3)
        for (int k = 0; k < 3; ++k) residual3[k] = sqrt(rr[k]);
    free(x); free(v); free(sg); free(b); free(r); free(d); free(Ad);
    return it;
}

void vpmo_nextstep(double *P, int64_t np, const vpmo_schemes *s, double dt, const double *Uinf,
                   int32_t relax, double *t, int64_t *nt) {
    const double zeta0 = vpmo_zeta(s->kernel, 0.0);
    if (np > 0) {
        if (s->integration == VPMO_INTEGRATION_EULER) {
            vpmo_field_sfs(P, np, s, 1.0, 1.0, *t, *nt);
            for (int64_t i = 0; i < np; ++i) {
                double *p = PCOL(P, i);
                if (is_static(p)) continue;
                double Msave[9];
                memcpy(Msave, p + VPMO_M, sizeof(Msave)); /* Euler keeps no q-storage */
                for (int k = 0; k < 9; ++k) p[VPMO_M + k] = 0;
                /* viscous update after relaxation in Euler: do the inviscid part first */
                vpmo_schemes sv = *s;
                sv.viscous = VPMO_VISCOUS_INVISCID;
                vpmo_update_particle(p, &sv, 0.0, 1.0, dt, Uinf, zeta0);
                if (relax) vpmo_relax_particle(p, s->relaxation, s->rlxf);
                if (s->viscous == VPMO_VISCOUS_CORESPREADING)
                    p[VPMO_SIGMA] = sqrt(p[VPMO_SIGMA] * p[VPMO_SIGMA] + 2 * s->nu * dt);
                memcpy(p + VPMO_M, Msave, sizeof(Msave));
            }
            if (s->viscous == VPMO_VISCOUS_CORESPREADING && s->cs_sgm0 > 0)
                vpmo_corespreading_reset(P, np, s->kernel, s->cs_sgm0, s->cs_beta, s->cs_itmax, s->cs_tol, NULL);
        } else {
            static const double AB[3][2] = {
                {0.0, 1.0 / 3.0}, {-5.0 / 9.0, 15.0 / 16.0}, {-153.0 / 128.0, 8.0 / 15.0}};
            for (int64_t i = 0; i < np; ++i) {
                double *p = PCOL(P, i);
                if (!is_static(p))
                    for (int k = 0; k < 9; ++k) p[VPMO_M + k] = 0;
            }
            for (int st = 0; st < 3; ++st) {
                double a = AB[st][0], b = AB[st][1];
                vpmo_field_sfs(P, np, s, a, b, *t, *nt);
                for (int64_t i = 0; i < np; ++i) {
                    double *p = PCOL(P, i);
                    if (!is_static(p)) vpmo_update_particle(p, s, a, b, dt, Uinf, zeta0);
                }
            }
            /* spatial adaptation is checked once the last substep is done (A.8), before the relaxation evaluation */
            if (s->viscous == VPMO_VISCOUS_CORESPREADING && s->cs_sgm0 > 0)
                vpmo_corespreading_reset(P, np, s->kernel, s->cs_sgm0, s->cs_beta, s->cs_itmax, s->cs_tol, NULL);
            if (relax && s->relaxation != VPMO_RELAX_NONE) {
                vpmo_field_uj(P, np, s, 1, 0, 0);
                for (int64_t i = 0; i < np; ++i) {
                    double *p = PCOL(P, i);
                    if (!is_static(p)) vpmo_relax_particle(p, s->relaxation, s->rlxf);
                }
            }
        }
    }
    *t += dt;
    *nt += 1;
}
