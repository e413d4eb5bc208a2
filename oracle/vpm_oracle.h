/*
 * vpm_oracle.h — CPU restatement (plain C + OpenMP) of the rVPM particle-field hot path that
 * FLOWUnsteady drives through FLOWVPM:  UJ_direct, Estr_direct, the dynamic SFS coefficient,
 * rungekutta3 / euler, Pedrizzetti relaxation and the core-spreading sigma update.
 *
 * THIS IS TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` legs may load it.  The product (flowunsteady_b200/) never links or calls it.
 *
 * PARITY UNPINNED: the reference's own implementation of this path is in the un-vendored Julia
 * dependency FLOWVPM (Project.toml:15,25 of the reference) and cannot be compiled or imported here
 * (no julia, no sources).  What IS pinned by in-tree reference code:
 *   - the Biot-Savart kernel form, the 1/(4 pi) literal, the r != 0 skip and the use of the SOURCE
 *     particle's sigma:  src/FLOWUnsteady_processing_force.jl:879-929 (_Ffv_direct);
 *   - the governing equations: docs/src/theory/rvpm.md:89-100 (U), :107-235 (dx/dt, dGamma/dt,
 *     dsigma/dt; f=0, g=1/5), :251-296 (E_str, C_d, clipping), :363-370 (schemes);
 *   - the call order: src/FLOWUnsteady_simulation.jl:339-447, :494-572.
 * Everything else follows the published FLOWVPM algorithm as restated in SURVEY.md Appendix A
 * (tagged UPSTREAM-RECALL there) and is validated against analytic identities and 50-digit mpmath
 * evaluations (tests/golden/, tests/test_oracle_*.py).
 *
 * Particle record: column-major matrix, one column of `VPMO_NFIELDS` doubles per particle
 * (the matrix addressing `pfield.particles[vpm.SIGMA_INDEX, i]` is confirmed at
 * src/FLOWUnsteady_simulation.jl:509-510).  0-based offsets below.
 */
#ifndef VPM_ORACLE_H
#define VPM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPMO_NFIELDS 43
#define VPMO_X 0      /* 0:3   position                              */
#define VPMO_GAMMA 3  /* 3:6   vectorial circulation                 */
#define VPMO_SIGMA 6  /* 6     core size                             */
#define VPMO_VOL 7    /* 7     volume                                */
#define VPMO_CIRC 8   /* 8     scalar circulation                    */
#define VPMO_U 9      /* 9:12  velocity                              */
#define VPMO_W 12     /* 12:15 vorticity                             */
#define VPMO_J 15     /* 15:24 J[i,j] = du_i/dx_j at 15 + i + 3 j    */
#define VPMO_PSE 24   /* 24:27 particle-strength-exchange scratch    */
#define VPMO_M 27     /* 27:36 scratch M[i,j] at 27 + i + 3 j        */
#define VPMO_C 36     /* 36:39 C_d, <Gamma.L>, <Gamma.m>             */
#define VPMO_SFS 39   /* 39:42 SFS term E_str                        */
#define VPMO_STATIC 42/* 42    static flag (>0: embedded particle)   */

/* kernel ids (names at src/FLOWUnsteady_simulation.jl:37) */
enum { VPMO_KERNEL_GAUSSIANERF = 0, VPMO_KERNEL_WINCKELMANS = 1, VPMO_KERNEL_GAUSSIAN = 2,
       VPMO_KERNEL_SINGULAR = 3 };
/* relaxation ids (src/FLOWUnsteady_simulation.jl:44) */
enum { VPMO_RELAX_NONE = 0, VPMO_RELAX_PEDRIZZETTI = 1, VPMO_RELAX_CORRECTEDPEDRIZZETTI = 2 };
/* SFS ids (src/FLOWUnsteady_simulation.jl:39; examples/rotorhover/rotorhover.jl:53-55,172-182) */
enum { VPMO_SFS_NONE = 0, VPMO_SFS_CONSTANT = 1, VPMO_SFS_DYNAMIC = 2 };
/* clipping / control bit masks */
enum { VPMO_CLIP_BACKSCATTER = 1 };
enum { VPMO_CTRL_DIRECTIONAL = 1, VPMO_CTRL_MAGNITUDE = 2, VPMO_CTRL_SIGMASENSOR = 4 };
/* viscous ids */
enum { VPMO_VISCOUS_INVISCID = 0, VPMO_VISCOUS_CORESPREADING = 1 };
/* integration ids */
enum { VPMO_INTEGRATION_EULER = 0, VPMO_INTEGRATION_RK3 = 1 };

typedef struct {
    int32_t kernel;        /* VPMO_KERNEL_*                                                     */
    double f, g;           /* formulation: rVPM (0, 1/5); cVPM (0, 0)   rvpm.md:197-239         */
    int32_t transposed;    /* 1: S = (Gamma . grad')u  (default, simulation.jl:41)              */
    int32_t relaxation;    /* VPMO_RELAX_*                                                      */
    double rlxf;           /* relaxation factor (0.3 upstream default)                          */
    int32_t sfs;           /* VPMO_SFS_*                                                        */
    double alpha;          /* test-filter ratio (0.999 two-level, 0.667 three-level)            */
    double sfs_rlxf;       /* Lagrangian-average relaxation (0.005, rotorhover.jl:179)          */
    double minC, maxC;     /* clamp of |C_d| (0, 1)                                             */
    double Cs;             /* ConstantSFS coefficient                                           */
    int32_t force_positive;/* pseudo3level_positive                                             */
    int32_t clippings;     /* VPMO_CLIP_* mask                                                  */
    int32_t controls;      /* VPMO_CTRL_* mask                                                  */
    int32_t viscous;       /* VPMO_VISCOUS_*                                                    */
    double nu;             /* kinematic viscosity for core spreading                            */
    int32_t integration;   /* VPMO_INTEGRATION_*                                                */
    double cs_sgm0;        /* CoreSpreading: reset core size (<= 0: never reset)   rotorhover.jl:170     */
    double cs_beta;        /*   reset when sigma/sgm0 > beta (1.5)                                       */
    int32_t cs_itmax;      /*   RBF conjugate-gradient iterations (15)                                   */
    double cs_tol;         /*   RBF residual tolerance (1e-3)                                            */
} vpmo_schemes;

/* Sets `s` to FLOWUnsteady's defaults (src/FLOWUnsteady_simulation.jl:36-44): rVPM, gaussianerf,
 * transposed, pedrizzetti(0.3), SFS_none, Inviscid, rungekutta3. */
void vpmo_default_schemes(vpmo_schemes *s);

/* Regularising function of kernel `kernel` at r_hat = r/sigma: *g = g(r_hat), *dg = g'(r_hat). */
void vpmo_g_dgdr(int32_t kernel, double r_hat, double *g, double *dg);
/* Radial basis zeta(r_hat); zeta_sigma(x) = zeta(|x|/sigma)/sigma^3. */
double vpmo_zeta(int32_t kernel, double r_hat);

/* UJ_direct over raw arrays.  Sources: xs[3*j..], gs[3*j..], sig[j].  Targets:
 * xt[3*i..].  Accumulates into U[3*i..] and J[9*i..] (J[i + 3 j] layout) — callers zero first.
 * accum = 0: plain double, sources in index order (the reference's form);
 * accum = 1: long-double accumulators (truth for the 1e-12 checks).                            */
void vpmo_uj_direct(int32_t kernel, int64_t ns, const double *xs, const double *gs, const double *sig,
                    int64_t nt, const double *xt, double *U, double *J, int32_t accum);

/* Estr_direct over raw arrays.  Js/Jt: 9 doubles per particle (J[i + 3 j]).
 * Accumulates into SFS[3*i..].                                                                 */
void vpmo_estr_direct(int32_t kernel, int32_t transposed, int64_t ns, const double *xs, const double *gs,
                      const double *sig, const double *Js, int64_t nt, const double *xt, const double *Jt,
                      double *SFS, int32_t accum);

/* The in-tree pin: restates _Ffv_direct (src/FLOWUnsteady_processing_force.jl:879-929).
 * sources = bound vortices b (x, Gamma, sigma), targets = free vortices f (x, Gamma).
 * M6[6*b..]: M[0:3] = sum_f U_b(x_f) x Gamma_f ; M[3:6] = sum_f g K(x_b - x_f) x (Gamma_b x Gamma_f). */
void vpmo_ffv_direct(int32_t kernel, int64_t nb, const double *xb, const double *gb, const double *sb,
                     int64_t nf, const double *xf, const double *gf, double *M6);

/* ---- field-level functions on the 43 x np column-major particle matrix -------------------- */

/* _reset_particles: U, J, PSE <- 0 (statics included).  _reset_particles_sfs: SFS <- 0.        */
void vpmo_reset_particles(double *P, int64_t np);
void vpmo_reset_particles_sfs(double *P, int64_t np);

/* pfield.UJ(pfield; reset, reset_sfs, sfs) with UJ_direct (+ Estr_direct when sfs != 0).        */
void vpmo_field_uj(double *P, int64_t np, const vpmo_schemes *s, int32_t reset, int32_t reset_sfs,
                   int32_t sfs);

/* pfield.SFS(pfield; a, b): evaluates U, J (and SFS, C) as the scheme requires.
 * `t`, `nt` are the field time and step count (control_magnitude estimates dt = t/nt).          */
void vpmo_field_sfs(double *P, int64_t np, const vpmo_schemes *s, double a, double b, double t, int64_t nt);

/* vpm.nextstep(pfield, dt; relax) (src/FLOWUnsteady_simulation.jl:358): one euler / rungekutta3
 * step including relaxation and core spreading (no RBF re-fit).  Uinf[3] is the freestream.
 * Advances *t and *nt.                                                                          */
void vpmo_nextstep(double *P, int64_t np, const vpmo_schemes *s, double dt, const double *Uinf,
                   int32_t relax, double *t, int64_t *nt);

/* zeta pass (vpm.zeta_direct / zeta_fmm): out[3*i..] (+)= sum_q v_q zeta_sigma_q(x_i - x_q), v = vs[3*q..].  With v = Gamma
 * this is the particle-approximated vorticity omega(x_i) (rvpm.md:80-86). */
void vpmo_zeta_direct(int32_t kernel, int64_t ns, const double *xs, const double *vs, const double *sig, int64_t nt,
                      const double *xt, double *out);

/* Core-spreading spatial adaptation (DOC rvpm.md:367; ctor examples/rotorhover/rotorhover.jl:170):
 * if any non-static sigma/sgm0 > beta: omega_p = sum_q Gamma_q zeta_sigma_q(x_p - x_q) with the SPREAD cores, then every
 * non-static sigma <- sgm0 and Gamma is re-fitted by conjugate gradients (per component, at most itmax iterations, stop when
 * every component's residual 2-norm < tol) so that sum_q Gamma_q zeta_sgm0(x_p - x_q) = omega_p.  Static particles keep
 * sigma and Gamma and act as fixed sources on both sides.  Returns the number of CG iterations (0: no reset needed). */
int32_t vpmo_corespreading_reset(double *P, int64_t np, int32_t kernel, double sgm0, double beta, int32_t itmax, double tol,
                                 double *residual3);

/* Per-particle pieces, exported so the tests can pin each one separately. */
void vpmo_relax_particle(double *p, int32_t relaxation, double rlxf);
void vpmo_update_particle(double *p, const vpmo_schemes *s, double a, double b, double dt,
                          const double *Uinf, double zeta0);

int32_t vpmo_num_threads(void);
void vpmo_set_num_threads(int32_t n);

/* fmm_oracle.c: the reference's FMM restated (ExaFMM-style solid harmonics of degree < p, dual tree traversal with
 * (R_i + R_j) < theta |c_i - c_j|, near field = vpmo_uj_direct).  U (n x 3), J (n x 9) at every particle from every particle. */
int32_t vpmo_fmm_uj(int32_t kernel, int32_t p, int32_t ncrit, double theta, double leaf_sigmas, int64_t n, const double *x,
                    const double *g, const double *sig, double *U, double *J, int64_t *stats);

#ifdef __cplusplus
}
#endif
#endif
