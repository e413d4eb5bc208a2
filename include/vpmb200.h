/*
 * vpmb200.h — C ABI of the B200-native rVPM particle-field engine (libvpmb200.so).
 *
 * This is the drop-in boundary for the hot path FLOWUnsteady drives every time step through the FLOWVPM
 * `ParticleField` plugin API.  Citations are into /root/reference (FLOWUnsteady v3.4.0):
 *
 *   reference call site                                              replaced by
 *   ---------------------------------------------------------------  ------------------------------------
 *   vpm.ParticleField(max, T; formulation, viscous, kernel, UJ, SFS,  vpmb200_create + vpmb200_set_schemes
 *       integration, transposed, relaxation, fmm)                     (src/FLOWUnsteady_simulation.jl:239-253)
 *   pfield.UJ(pfield)               (simulation.jl:544,               vpmb200_uj
 *                                    processing_force.jl:238)
 *   vpm._reset_particles(pfield)    (processing_force.jl:237)         vpmb200_reset_particles
 *   pfield.SFS(pfield; a, b)        (inside vpm.nextstep)             vpmb200_sfs
 *   vpm.nextstep(pfield, dt; relax) (simulation.jl:358)               vpmb200_nextstep
 *   probes appended + pfield.UJ     (simulation.jl:494-570)           vpmb200_uj_probe
 *   pfield.particles[:, 1:np]       (simulation.jl:509-510)           vpmb200_upload / vpmb200_download
 *   vpm.add_particle / remove_particle (simulation.jl:363,486,551)    vpmb200_add_particles / vpmb200_remove_particle
 *
 * Conventions
 *   - All functions return 0 on success, a negative VPMB200_E* code otherwise; no exceptions cross the ABI.
 *     vpmb200_last_error(h) returns a human-readable message for the last failure on that handle.
 *   - Host pointers are borrowed for the duration of the call only.  The engine owns all device memory.
 *   - The particle matrix is the reference's: column-major, `nfields` (= 43) doubles per particle, one column per
 *     particle, leading dimension `ld` >= nfields (field offsets below).
 *   - A handle is bound to one CUDA device and one stream; calls on a handle are serialised by the caller.
 *   - There is NO CPU fallback: create fails with VPMB200_ENODEVICE when no sm_100 GPU is present.
 */
#ifndef VPMB200_H
#define VPMB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VPMB200_NFIELDS 43
/* 0-based row offsets in a particle column */
#define VPMB200_X 0
#define VPMB200_GAMMA 3
#define VPMB200_SIGMA 6
#define VPMB200_VOL 7
#define VPMB200_CIRCULATION 8
#define VPMB200_U 9
#define VPMB200_VORTICITY 12
#define VPMB200_J 15 /* J[i,j] = du_i/dx_j at 15 + i + 3 j */
#define VPMB200_PSE 24
#define VPMB200_M 27
#define VPMB200_C 36
#define VPMB200_SFS 39
#define VPMB200_STATIC 42

/* field-group bit masks for vpmb200_upload / vpmb200_download */
#define VPMB200_FM_X (1u << 0)
#define VPMB200_FM_GAMMA (1u << 1)
#define VPMB200_FM_SIGMA (1u << 2)
#define VPMB200_FM_VOL (1u << 3)
#define VPMB200_FM_CIRCULATION (1u << 4)
#define VPMB200_FM_U (1u << 5)
#define VPMB200_FM_VORTICITY (1u << 6)
#define VPMB200_FM_J (1u << 7)
#define VPMB200_FM_PSE (1u << 8)
#define VPMB200_FM_M (1u << 9)
#define VPMB200_FM_C (1u << 10)
#define VPMB200_FM_SFS (1u << 11)
#define VPMB200_FM_STATIC (1u << 12)
#define VPMB200_FM_ALL 0x1fffu
/* what the host owns between calls (everything the integrator reads) */
#define VPMB200_FM_STATE (VPMB200_FM_X | VPMB200_FM_GAMMA | VPMB200_FM_SIGMA | VPMB200_FM_VOL | \
                          VPMB200_FM_CIRCULATION | VPMB200_FM_C | VPMB200_FM_STATIC)

/* error codes */
#define VPMB200_OK 0
#define VPMB200_EINVAL (-1)
#define VPMB200_ENODEVICE (-2)
#define VPMB200_ECUDA (-3)
#define VPMB200_ECAPACITY (-4) /* more particles than max_particles (reference: simulation.jl:255-261) */
#define VPMB200_ENOTSUP (-5)

/* scheme ids — the reference's scheme objects (src/FLOWUnsteady_simulation.jl:36-44) */
enum { VPMB200_KERNEL_GAUSSIANERF = 0, VPMB200_KERNEL_WINCKELMANS = 1, VPMB200_KERNEL_GAUSSIAN = 2,
       VPMB200_KERNEL_SINGULAR = 3 };
enum { VPMB200_RELAX_NONE = 0, VPMB200_RELAX_PEDRIZZETTI = 1, VPMB200_RELAX_CORRECTEDPEDRIZZETTI = 2 };
enum { VPMB200_SFS_NONE = 0, VPMB200_SFS_CONSTANT = 1, VPMB200_SFS_DYNAMIC = 2 };
enum { VPMB200_CLIP_BACKSCATTER = 1 };
enum { VPMB200_CTRL_DIRECTIONAL = 1, VPMB200_CTRL_MAGNITUDE = 2 };
enum { VPMB200_VISCOUS_INVISCID = 0, VPMB200_VISCOUS_CORESPREADING = 1 };
enum { VPMB200_INTEGRATION_EULER = 0, VPMB200_INTEGRATION_RK3 = 1 };
enum { VPMB200_UJ_DIRECT = 0, VPMB200_UJ_FMM = 1 };

/* Mirrors the scheme fields of vpm.ParticleField (simulation.jl:239-244).  Same layout as oracle's vpmo_schemes
 * for the first 22 members so tests can drive both from one description. */
typedef struct {
    int32_t kernel;         /* vpm_kernel: gaussianerf | winckelmans | gaussian | singular                */
    double f, g;            /* vpm_formulation: rVPM (0, 1/5) | cVPM (0, 0)                                */
    int32_t transposed;     /* vpm_transposed                                                              */
    int32_t relaxation;     /* vpm_relaxation                                                              */
    double rlxf;            /*   relaxation factor (0.3)                                                   */
    int32_t sfs;            /* vpm_SFS: none | constant | dynamic                                          */
    double alpha;           /*   DynamicSFS alpha (test-filter ratio)                                      */
    double sfs_rlxf;        /*   DynamicSFS rlxf (Lagrangian average)                                      */
    double minC, maxC;      /*   DynamicSFS clamps                                                         */
    double Cs;              /*   ConstantSFS coefficient                                                   */
    int32_t force_positive; /*   pseudo3level_positive                                                     */
    int32_t clippings;      /*   VPMB200_CLIP_* mask                                                       */
    int32_t controls;       /*   VPMB200_CTRL_* mask                                                       */
    int32_t viscous;        /* vpm_viscous: Inviscid | CoreSpreading                                            */
    double nu;              /*   kinematic viscosity                                                       */
    int32_t integration;    /* vpm_integration: euler | rungekutta3                                        */
    double cs_sgm0;         /* CoreSpreading(nu, sgm0, zeta; beta, itmax, tol): reset core size, <= 0 = never reset */
    double cs_beta;         /*   reset when sigma/sgm0 > beta (1.5)                                        */
    int32_t cs_itmax;       /*   RBF conjugate-gradient iterations (15)                                    */
    double cs_tol;          /*   RBF residual tolerance (1e-3)                                             */
    /* --- beyond the oracle's struct --- */
    int32_t uj;             /* vpm_UJ: direct | fmm                                                        */
    int32_t fmm_p;          /* vpm_fmm = vpm.FMM(; p=4, ncrit=50, theta=0.4, nonzero_sigma) simulation.jl:43 */
    int32_t fmm_ncrit;
    double fmm_theta;
    int32_t fmm_nonzero_sigma; /* 0 = false; 1 = true (far field only beyond 5 core sizes of clearance between two cells'
                                  closest points); k >= 2 = true with k core sizes                                */
} vpmb200_schemes;

typedef struct vpmb200_engine* vpmb200_handle;

/* FLOWUnsteady's defaults (simulation.jl:36-44, except uj = direct). */
int32_t vpmb200_default_schemes(vpmb200_schemes* s);

/* max_particles: capacity (reference: max_particles, simulation.jl:124,239).  nfields must be 43.
 * float_bits: 64 (FP64 pair arithmetic) or 32 (FP32 pair arithmetic, FP64 state; reference knob vpm_floattype,
 * simulation.jl:137).  device: CUDA ordinal. */
int32_t vpmb200_create(int64_t max_particles, int32_t nfields, int32_t float_bits, int32_t device,
                       vpmb200_handle* out);
int32_t vpmb200_destroy(vpmb200_handle h);
const char* vpmb200_last_error(vpmb200_handle h);

int32_t vpmb200_set_schemes(vpmb200_handle h, const vpmb200_schemes* s);
int32_t vpmb200_get_schemes(vpmb200_handle h, vpmb200_schemes* s);

/* Field time and step counter (pfield.t, pfield.nt). */
int32_t vpmb200_set_time(vpmb200_handle h, double t, int64_t nt);
int32_t vpmb200_get_time(vpmb200_handle h, double* t, int64_t* nt);

/* Replace the device field by columns 0..np-1 of `particles` (only the groups in field_mask are copied; others
 * keep their device values).  np becomes the particle count. */
int32_t vpmb200_upload(vpmb200_handle h, const double* particles, int64_t ld, int64_t np, uint32_t field_mask);
/* Copy the groups in field_mask of particles 0..np-1 back into `particles`. */
int32_t vpmb200_download(vpmb200_handle h, double* particles, int64_t ld, int64_t np, uint32_t field_mask);
int32_t vpmb200_get_np(vpmb200_handle h, int64_t* np);
/* Page-lock (cudaHostRegister) / release the caller's particle matrix so that vpmb200_upload / vpmb200_download move it by
 * DMA at full PCIe rate straight from / into it; a pageable matrix (what `Matrix{Float64}(undef, 43, max)` is) still works
 * but is staged by the driver at a few GB/s.  Call once after allocating `pfield.particles`
 * (/root/reference/src/FLOWUnsteady_simulation.jl:239-253) and unregister before it is freed (the Julia stub does both in
 * `_handle` and its finalizer).  No handle: errors are read with vpmb200_last_error(NULL). */
int32_t vpmb200_host_register(void* ptr, uint64_t bytes);
int32_t vpmb200_host_unregister(void* ptr);

/* vpm.add_particle: append n columns.  vpm.remove_particle(pfield, i): 0-based i; the last particle is swapped
 * into slot i (the reference's behaviour). */
int32_t vpmb200_add_particles(vpmb200_handle h, const double* cols, int64_t ld, int64_t n);
int32_t vpmb200_remove_particle(vpmb200_handle h, int64_t i);

/* Wake treatments — FLOWUnsteady's remove_particles_strength / _sigma / _box / _sphere runtime functions
 * (src/FLOWUnsteady_processing.jl:50-187) as one device-side compaction.  The surviving particles end up in exactly the
 * order the reference's loop (i = np..1, vpm.remove_particle(i) = move the last particle into slot i) produces.
 *   STRENGTH: params = { minGamma2, maxGamma2 }                      keep iff min <= |Gamma|^2 <= max
 *   SIGMA   : params = { minsigma, maxsigma }                        keep iff min <= sigma <= max
 *   BOX     : params = { Pmin[3], Pmax[3], O[3] }                    remove if X - O lies outside the box
 *   SPHERE  : params = { Rsphere2, centre[3] }                       remove if |X - centre|^2 > Rsphere2
 * *removed receives the number of particles removed. */
enum { VPMB200_REMOVE_STRENGTH = 1, VPMB200_REMOVE_SIGMA = 2, VPMB200_REMOVE_BOX = 3, VPMB200_REMOVE_SPHERE = 4 };
int32_t vpmb200_remove_where(vpmb200_handle h, int32_t criterion, const double* params, int64_t* removed);

/* vpm.zeta_direct / zeta_fmm: W (rows 12:15) <- sum_q Gamma_q zeta_sigma_q(x_p - x_q), the particle-approximated vorticity. */
int32_t vpmb200_zeta(vpmb200_handle h);
/* CoreSpreading's spatial adaptation: if any non-static sigma/sgm0 > beta, store omega (spread cores) in
 * W, set sigma <- sgm0 and re-fit Gamma by conjugate gradients.  *iters: CG iterations done (0 = no reset was needed);
 * residual3: final residual 2-norm per component (may be NULL).  vpmb200_nextstep calls this itself when
 * viscous = CoreSpreading and cs_sgm0 > 0. */
int32_t vpmb200_corespreading_reset(vpmb200_handle h, int32_t* iters, double* residual3);

/* Monitors (vpm.monitor_enstrophy, vpm.monitor_Cd; src/FLOWUnsteady_monitors.jl:614,697): out[0] = enstrophy
 * 0.5 sum Gamma.omega (omega = curl u from J), out[1] = mean C_d over particles with C_d != 0, out[2] = its standard
 * deviation, out[3] = number of particles with C_d != 0, out[4] = number of static particles, out[5] = sum |Gamma|. */
int32_t vpmb200_monitors(vpmb200_handle h, double* out6);

/* vpm._reset_particles: U, J, PSE <- 0.   _reset_particles_sfs: SFS <- 0. */
int32_t vpmb200_reset_particles(vpmb200_handle h);
int32_t vpmb200_reset_particles_sfs(vpmb200_handle h);

/* pfield.UJ(pfield; reset, reset_sfs, sfs): U and J at every particle from every particle; with sfs != 0 also
 * accumulates the SFS term E_str.  NOTE the reference call `pfield.UJ(pfield)` is reset=1, reset_sfs=0, sfs=0. */
int32_t vpmb200_uj(vpmb200_handle h, int32_t reset, int32_t reset_sfs, int32_t sfs);

/* Velocity (and optionally J) induced by the field at m probe points X (3 x m column-major).  Equivalent to the
 * reference's add_probe + pfield.UJ + get_U sequence (simulation.jl:536-547) without evaluating the other
 * targets.  U: 3 x m.  J: 9 x m or NULL. */
int32_t vpmb200_uj_probe(vpmb200_handle h, const double* X, int64_t m, double* U, double* J);

/* The same with Vvpm_on_Xs's two options (simulation.jl:494-570):
 *   fsgm   — `sigmafactor_vpmonvlm`.  The reference's loop (simulation.jl:507-513, undone at :555-561) indexes the particle
 *            matrix with ONE subscript, so it multiplies sigma of the FIRST particle by fsgm once per static particle present
 *            in the field when it runs (normally none — the statics are added after it — which makes the option a no-op in
 *            the reference).  Reproduced literally, including the k sequential roundings of *= and /=.
 *   mirror — method of images in the plane set with vpmb200_set_mirror: images of every source (field, statics, and the images
 *            vpmb200_set_mirror(enabled) already added) join the sources for this call (simulation.jl:518-535).
 * The static set parked with vpmb200_set_statics takes the place of `static_particles_fun(pfield, t, dt)`. */
int32_t vpmb200_uj_probe_ex(vpmb200_handle h, const double* X, int64_t m, double fsgm, int32_t mirror, double* U, double* J);

/* Static-particle fast path.  FLOWUnsteady appends the embedded (static) particles to the field before every nextstep and
 * removes them after it (simulation.jl:355-365), and again around every probe evaluation (:515).  vpmb200_set_statics parks
 * the n columns `cols` (leading dimension ld; their static flag is forced to 1) device-side BEHIND the field instead:
 * vpmb200_nextstep (which consumes the set), vpmb200_uj and vpmb200_uj_probe[_ex] treat them as part of the field while the
 * field's step counter nt equals `generation` — a set left over from another step is ignored.  Any call that changes the
 * field's particles (upload, add, remove) drops the set.  Equivalent to add_particle x n -> call -> remove_particle x n. */
int32_t vpmb200_set_statics(vpmb200_handle h, const double* cols, int64_t ld, int64_t n, int64_t generation);
int32_t vpmb200_get_statics(vpmb200_handle h, int64_t* n, int64_t* generation);
/* Method of images (run_simulation's mirror, mirror_X, mirror_normal; vehicle_vlm_unsteady.jl:245-260): while enabled, the
 * image of EVERY particle (field + static set) joins the static set wherever that set is used, built on the device with the
 * reference's expressions  Xm = X - dot(2 (X - X0), n) n,  Gm = 2 dot(G, n) G / |G| - G  (sic).  Needs max_particles >= twice
 * the particle count.  The plane is stored even when enabled = 0 (vpmb200_uj_probe_ex's mirror flag uses it). */
int32_t vpmb200_set_mirror(vpmb200_handle h, int32_t enabled, const double* X0, const double* normal);

/* pfield.SFS(pfield; a, b): evaluates U, J (+ SFS, C_d) as the SFS scheme prescribes for RK coefficients (a, b). */
int32_t vpmb200_sfs(vpmb200_handle h, double a, double b);

/* vpm.nextstep(pfield, dt; relax): one euler / rungekutta3 step (UJ + SFS + update + relaxation + core
 * spreading); Uinf is the freestream pfield.Uinf(t) evaluated by the host (3 doubles). */
int32_t vpmb200_nextstep(vpmb200_handle h, double dt, const double* Uinf, int32_t relax);

/* Monitors: np, number of non-finite X/Gamma/sigma entries, enstrophy 0.5 sum Gamma.omega is left to the host. */
int32_t vpmb200_count_nonfinite(vpmb200_handle h, int64_t* count);

/* ---- device-level hooks (multi-GPU sharding, benchmarking; pointers are CUDA device pointers) ------------- */

/* Pointer to SoA field row `field` (0..42) of the device state: element i of that row is particle i. */
int32_t vpmb200_device_field(vpmb200_handle h, int32_t field, double** ptr, int64_t* ld);
/* The CUDA stream (cudaStream_t) every call on this handle is enqueued on. */
int32_t vpmb200_stream(vpmb200_handle h, void** stream);
/* Engine options.  "direct_sort" (default 1): visit targets and source tiles of the DIRECT path in Morton order
 * internally (results are returned in particle order); 0 keeps the caller's particle order.
 * "fmm_table_copies" (1 or 8, default 8): shared-memory layout of the Gaussian-erf table in the UJ_fmm near-field kernel;
 * 8 = one copy per 16-byte bank group (conflict-free lookups, one 16-warp CTA per SM), 1 = single copy (two 8-warp CTAs).
 * Results do not depend on it. */
int32_t vpmb200_set_option(vpmb200_handle h, const char* name, int64_t value);
/* Tree statistics of the last UJ_fmm evaluation: stats[0..4] = cells, leaves, levels, M2L pairs, P2P (leaf) pairs. */
int32_t vpmb200_fmm_stats(vpmb200_handle h, int64_t* stats);
/* Instrumentation of the DIRECT path for the current field (what bench.py reports as `tile_far_fraction`): stats[0] = target
 * blocks (256 targets each), stats[1] = source tiles (256 sources each), stats[2] = (block, tile) pairs whose boxes are
 * farther apart than the tile's T_FAR sigma_max (K1 runs its branch-free singular loop on them, K2 skips them),
 * stats[3] = all (block, tile) pairs.  Uses the same ordering (direct_sort) and boxes as vpmb200_uj. */
int32_t vpmb200_direct_tile_stats(vpmb200_handle h, int64_t* stats);
/* Device time (ms, CUDA events on the handle's stream) of the sections of the LAST UJ_fmm evaluation: ms6 = { sort + tree build,
 * interaction lists (traversal sweeps with their read-backs + list sorts), upward pass, M2L + L2L, L2P + near field, E_str near
 * field }.  Diagnostics (load balance across ranks, bench.py). */
int32_t vpmb200_fmm_times(vpmb200_handle h, double* ms6);
/* Number of CUDA kernels this handle has enqueued since creation (bench.py reports the per-step delta). */
int32_t vpmb200_launch_count(vpmb200_handle h, uint64_t* count);
/* Block the host until all enqueued work on the handle has finished. */
int32_t vpmb200_synchronize(vpmb200_handle h);

/* Source tiles: what the pair kernels stream.  A tile is 256 source records of 10 doubles + one 10-double header
 * (bounding box, far-field radius), 2570 doubles in all; n particles pack into ceil(n/256) tiles.  Ranks all-gather
 * whole tiles (NCCL) to shard the direct path, so tile sets from different ranks simply concatenate. */
int64_t vpmb200_tiles_for(int64_t nparticles);      /* number of tiles for n particles                       */
int64_t vpmb200_tile_doubles(void);                 /* doubles per tile (2570)                               */
/* Pack the LOCAL particles [0, np) into dst (device pointer, >= tiles_for(np) * tile_doubles() doubles). */
int32_t vpmb200_pack_uj_records(vpmb200_handle h, double* dst);
int32_t vpmb200_pack_estr_records(vpmb200_handle h, double* dst);
/* U, J (or SFS) of the LOCAL particles from `ntiles` external source tiles (device pointer); accumulate != 0
 * adds to the current U, J.  The E_str variant always accumulates into SFS. */
int32_t vpmb200_uj_from_records(vpmb200_handle h, const double* tiles, int64_t ntiles, int32_t accumulate);
int32_t vpmb200_estr_from_records(vpmb200_handle h, const double* tiles, int64_t ntiles);
/* Multi-GPU UJ_fmm building block.  G (device) is a 24-row "mini-state" of ALL ntot particles of the job, row r of
 * particle i at G[r * ldg + i] with the particle record's row numbering (X 0:3, Gamma 3:6, sigma 6 filled by the caller;
 * U 9:12 and J 15:24 written by pass 0; rows 12:15 receive E_str in pass 1, which reads the J rows).  Every rank builds
 * the same tree and evaluates leaves [part, part+1) / nparts; rows of particles outside its share are written as zeros,
 * so the ranks' outputs combine with one all-reduce (flowunsteady_b200/dist.py). */
int32_t vpmb200_fmm_global(vpmb200_handle h, double* G, int64_t ldg, int64_t ntot, int32_t part, int32_t nparts, int32_t pass);

/* ---- multi-GPU UJ_fmm with a LOCAL ESSENTIAL TREE: per-rank phases (flowunsteady_b200/csrc/fmm_let.cuh) -----------------
 * One evaluation = bounds -> [all-reduce min/max] -> keys -> [all-reduce histogram (sum), bin sigma (max)] -> partition ->
 * pack -> [all-to-all of 7-double rows] -> build -> [all-gather of cells / multipoles / records] -> attach_tree +
 * attach_records -> evaluate -> [inverse all-to-all of 12-double rows] -> finish; with sfs: estr_records -> [all-gather of
 * records] -> attach_records -> estr_evaluate -> [inverse all-to-all of 3-double rows] -> finish(what = 1).
 * The bracketed collectives are the caller's (flowunsteady_b200/dist.py: NCCL); all pointers below are DEVICE pointers
 * except lohi6, send_counts, info4, ncells, nparticles and ptrs3 (host).  Results equal the one-GPU UJ_fmm to round-off. */
int32_t vpmb200_let_cell_bytes(void);                               /* bytes of one tree cell in the skeleton exchange     */
int32_t vpmb200_let_bounds(vpmb200_handle h, double* lohi6);        /* min xyz, max xyz of the local particles             */
/* Morton keys + sort of the local ("home") particles in the cube of the GLOBAL bounds, level-Lc histogram: *hist_dev = int32
 * [8^Lc] (all-reduce SUM in place), *binmax_dev = double [8^Lc] or NULL (all-reduce MAX in place; nonzero_sigma only). */
int32_t vpmb200_let_keys(vpmb200_handle h, const double* lohi6_global, int32_t Lc, void** hist_dev, void** binmax_dev);
/* Cuts the Morton curve into nparts key ranges at top-tree unit boundaries; send_counts[k] = local particles owned by rank k
 * (contiguous in the packed order).  use_work = 0: equal particle counts.  use_work != 0: equal WORK — vpmb200_let_evaluate
 * counts the interaction work of every level-Lc bin (near-field particle pairs, M2L translations; exact integer sums) into
 * the int64 [8^Lc] device array vpmb200_let_work returns; the caller all-reduces it (SUM, in place) after the evaluation and
 * the NEXT partition weighs every unit by it (every rank must pass the same flag). */
int32_t vpmb200_let_work(vpmb200_handle h, void** work_dev);
/* The cut itself, host arithmetic only (no device, no handle; what vpmb200_let_partition applies to the all-reduced arrays):
 * hist = particle count of every level-Lc Morton bin (int32 [8^Lc]), work = the counted work per bin or NULL (cut by count);
 * splitters[0 .. nparts] receives the key bounds, rank k owning Morton keys in [splitters[k], splitters[k + 1]).  A bound never
 * falls inside a cell of the global tree's top that holds <= ncrit particles (that cell is a leaf with one owner). */
int32_t vpmb200_let_cut(const int32_t* hist, const int64_t* work, int32_t Lc, int32_t ncrit, int32_t nparts, uint64_t* splitters);
int32_t vpmb200_let_partition(vpmb200_handle h, int32_t nparts, int32_t part, int32_t use_work, int64_t* send_counts);
int32_t vpmb200_let_pack(vpmb200_handle h, double* rows);           /* np rows of (x, y, z, Gamma, sigma), Morton order    */
/* Owner side: sort the n_own received rows, build this rank's part of the global octree, upward pass.  n_all = particles of
 * all ranks.  reuse != 0: same positions / strengths as the previous evaluation (DynamicSFS's second filter): only the
 * records are refreshed.  info4 = { cells, leaves, doubles per cell multipole, n_own }. */
int32_t vpmb200_let_build(vpmb200_handle h, const double* rows, int64_t n_own, int64_t n_all, int32_t reuse, int64_t* info4);
int32_t vpmb200_let_ptrs(vpmb200_handle h, void** ptrs3);           /* own cells, multipoles, records (send buffers)        */
/* all-gathered blocks (rank q at q * slot): tree skeletons + multipoles, then source records (UJ or E_str flavour) */
int32_t vpmb200_let_attach_tree(vpmb200_handle h, const void* cells_recv, const double* M_recv, int64_t slot_cells,
                                const int64_t* ncells, const int64_t* nparticles);
int32_t vpmb200_let_attach_records(vpmb200_handle h, const double* rec_recv, int64_t slot_n, const int64_t* nparticles);
/* Demand-driven halo (the memory-lean variant: a rank never holds another rank's multipoles or records unless its own
 * traversal needs them).  Only the SKELETONS are all-gathered -> attach_skeleton (nleaves = leaves of every rank's tree,
 * info4[1] of let_build) -> let_evaluate(stage 5: interaction lists) -> halo_plan: counts3[3 q + 0 / 1 / 2] = cells / leaves /
 * records this rank requests from rank q, *req_cells_dev = int32 owner-local cell ids, *req_leaf_dev = int32 pairs (first
 * record, count) in the owner's order, both grouped by owner -> [all-to-all of the requests] -> halo_serve on the owner
 * (requested multipoles, requested leaves' records back to back; M_out = NULL re-serves records only) -> [all-to-all of the
 * replies] -> halo_set(received multipoles, received records) -> let_evaluate(stage 6: M2L + L2L) -> let_evaluate(stage 2). */
int32_t vpmb200_let_attach_skeleton(vpmb200_handle h, const void* cells_recv, int64_t slot_cells, const int64_t* ncells,
                                    const int64_t* nparticles, const int64_t* nleaves);
int32_t vpmb200_let_halo_plan(vpmb200_handle h, int64_t* counts3, void** req_cells_dev, void** req_leaf_dev);
int32_t vpmb200_let_halo_serve(vpmb200_handle h, const void* req_cells, int64_t ncell, const void* req_leaf, int64_t nleaf,
                               double* M_out, double* rec_out);
int32_t vpmb200_let_halo_set(vpmb200_handle h, const double* M2, const double* rec2);
/* n_own rows of U(3), J(9) in arrival order.  stage 0: everything.  stage 1: interaction lists + far field only (needs
 * attach_tree, not the records — the record exchange can overlap it); stage 2: L2P + near field + the output rows;
 * stage 5: interaction lists only; stage 6: M2L + L2L only (the halo variant plans its requests between the two). */
int32_t vpmb200_let_evaluate(vpmb200_handle h, double* out_rows, int32_t reuse, int32_t stage);
int32_t vpmb200_let_estr_records(vpmb200_handle h);                 /* own records -> E_str flavour (needs evaluate's J)    */
int32_t vpmb200_let_estr_evaluate(vpmb200_handle h, double* out_rows);             /* n_own rows of E_str(3)                  */
/* Home side: res_rows in the packed order.  what = 0: U, J rows (reset != 0 overwrites and zeroes PSE, else accumulates);
 * what = 1: SFS rows += E_str. */
int32_t vpmb200_let_finish(vpmb200_handle h, const double* res_rows, int32_t what, int32_t reset);

/* The per-particle stages of pfield.SFS / nextstep, exposed so a sharded driver can interleave its exchange:
 * stage ids in vpmb200_stage. */
enum {
    VPMB200_STAGE_SCALE_SIGMA_TEST = 1,   /* sigma *= alpha (non-static)                 */
    VPMB200_STAGE_STORE_TEST = 2,         /* M[:,1] = S, M[:,2] = SFS                    */
    VPMB200_STAGE_SCALE_SIGMA_DOMAIN = 3, /* sigma /= alpha                              */
    VPMB200_STAGE_DYNAMIC_COEFF = 4,      /* M -= ..., C_d, clamp, flush M               */
    VPMB200_STAGE_CONSTANT_COEFF = 5,     /* C[1] = Cs                                   */
    VPMB200_STAGE_CLIP_CONTROL = 6,       /* clippings + controls                        */
    VPMB200_STAGE_ZERO_M = 7,             /* rungekutta3 q-storage reset                 */
    VPMB200_STAGE_UPDATE = 8,             /* one (a, b) substep incl. core spreading     */
    VPMB200_STAGE_RELAX = 9,              /* relaxation(p)                               */
    VPMB200_STAGE_UPDATE_EULER_RELAX = 10 /* euler update with relaxation(p) applied inline */
};
int32_t vpmb200_stage(vpmb200_handle h, int32_t stage, double a, double b, double dt, const double* Uinf);

/* ---- several GPUs behind ONE handle and ONE host thread (flowunsteady_b200/csrc/multi.inl) ---------------------------------
 * What a single-process host such as FLOWUnsteady's Julia loop (simulation.jl:339-447) needs to reach the 8 GPUs of a box with
 * the same upload / uj / nextstep / add / remove / download sequence.  Particles are sharded over `ngpus` per-device engines;
 * the GLOBAL particle order (the order of the host's matrix, on which vpm.remove_particle's swap-with-last is defined) is kept
 * as an index map on the host.  Direct U/J (+ E_str): source tiles are exchanged device-to-device (cudaMemcpyPeerAsync over
 * NVLink / NVSwitch, ordered by CUDA events; no NCCL, no host staging) and every device's pair kernels run concurrently;
 * per-particle stages are shard-local.  vpm_UJ = UJ_fmm runs the local-essential-tree phases (vpmb200_let_*) with the exchanges
 * done by peer copies (the sequence flowunsteady_b200/dist.py runs across processes with NCCL).  devices = NULL means 0 .. ngpus-1; the
 * same ordinal may appear more than once (several shards on one GPU: how the tests run on a one-GPU box). */
typedef struct vpmb200_multi* vpmb200_multi_handle;
int32_t vpmb200_multi_create(int64_t max_particles, int32_t nfields, int32_t float_bits, int32_t ngpus, const int32_t* devices,
                             vpmb200_multi_handle* out);
int32_t vpmb200_multi_destroy(vpmb200_multi_handle h);
const char* vpmb200_multi_last_error(vpmb200_multi_handle h);
int32_t vpmb200_multi_set_schemes(vpmb200_multi_handle h, const vpmb200_schemes* s);
int32_t vpmb200_multi_set_time(vpmb200_multi_handle h, double t, int64_t nt);
int32_t vpmb200_multi_get_time(vpmb200_multi_handle h, double* t, int64_t* nt);
int32_t vpmb200_multi_get_np(vpmb200_multi_handle h, int64_t* np);
int32_t vpmb200_multi_shard_sizes(vpmb200_multi_handle h, int64_t* n_per_gpu);
/* upload with all groups (or a new np) re-partitions the columns into ngpus contiguous blocks; a partial mask with the
 * current np refreshes those groups in place.  download needs np == the field's count. */
int32_t vpmb200_multi_upload(vpmb200_multi_handle h, const double* particles, int64_t ld, int64_t np, uint32_t field_mask);
int32_t vpmb200_multi_download(vpmb200_multi_handle h, double* particles, int64_t ld, int64_t np, uint32_t field_mask);
/* vpm.add_particle (simulation.jl:486): new global indices np, np+1, ...; stored on the least-loaded shard.
 * vpm.remove_particle(i) / the wake treatments: the reference's resulting GLOBAL order (swap with the last; the removal loop of
 * src/FLOWUnsteady_processing.jl:50-187), whatever the sharding. */
int32_t vpmb200_multi_add_particles(vpmb200_multi_handle h, const double* cols, int64_t ld, int64_t n);
int32_t vpmb200_multi_remove_particle(vpmb200_multi_handle h, int64_t i);
int32_t vpmb200_multi_remove_where(vpmb200_multi_handle h, int32_t criterion, const double* params, int64_t* removed);
/* Periodic rebalance: tails of the fullest shards move to the emptiest (device to device) until max - min <= tolerance x mean. */
int32_t vpmb200_multi_rebalance(vpmb200_multi_handle h, double tolerance, int64_t* moved);
/* vpmb200_set_statics on the sharded field: the step's static columns join the field (as its LAST global indices, on the
 * least-loaded shard, static flag forced) while the step counter equals `generation`; vpmb200_multi_nextstep removes them
 * again, any other mutation / transfer drops them first, vpmb200_multi_get_np does not count them. */
int32_t vpmb200_multi_set_statics(vpmb200_multi_handle h, const double* cols, int64_t ld, int64_t n, int64_t generation);
int32_t vpmb200_multi_uj(vpmb200_multi_handle h, int32_t reset, int32_t reset_sfs, int32_t sfs);
int32_t vpmb200_multi_sfs(vpmb200_multi_handle h, double a, double b);
int32_t vpmb200_multi_nextstep(vpmb200_multi_handle h, double dt, const double* Uinf, int32_t relax);
int32_t vpmb200_multi_uj_probe(vpmb200_multi_handle h, const double* X, int64_t m, double* U, double* J);
int32_t vpmb200_multi_synchronize(vpmb200_multi_handle h);
int32_t vpmb200_multi_engine(vpmb200_multi_handle h, int32_t k, vpmb200_handle* shard);   /* shard k's engine (borrowed)        */

/* Measurement helper: FP64 FMA peak of `device` in TFLOP/s (best of `repeats` launches of a register-only DFMA
 * kernel, `iters` x 128 FMAs per thread).  bench.py uses it as the roofline denominator of the FP64-bound pair
 * kernels (MEASURED_PEAKS.json has no FP64 entry). */
int32_t vpmb200_measure_fp64_peak(int32_t device, int32_t iters, int32_t repeats, double* tflops, double* ms);
/* The same with its evidence: out8 = { best measured DFMA TFLOP/s over several chain/CTA shapes (bench.py's roofline
 * denominator; ncu: 99.97 % FP64 pipe active), its launch ms, a clock64 diagnostic, the nominal FP64 PIPE rate
 * SMs x 64 lanes x 2 flop x nominal max SM clock (what ncu's sm__inst_executed_pipe_fp64 counts against),
 * measured / nominal, index of the best shape, SM count, the driver's nominal max SM clock (MHz) }. */
int32_t vpmb200_measure_fp64_peak2(int32_t device, int32_t iters, int32_t repeats, double* out8);

/* Library identification. */
const char* vpmb200_version(void);

#ifdef __cplusplus
}
#endif
#endif
