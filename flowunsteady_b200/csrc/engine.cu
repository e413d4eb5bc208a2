// engine.cu — C ABI (include/vpmb200.h) of the B200-native rVPM particle-field engine.
//
// One engine = one CUDA device + one stream + a device-resident SoA particle field.  Every entry point mirrors a
// call FLOWUnsteady makes on FLOWVPM's ParticleField (citations in include/vpmb200.h).  No CPU fallback exists:
// creation fails without an sm_100 device.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/vpmb200.h"
#include "estr_direct.cuh"
#include "field_kernels.cuh"
#include "fmm_host.cuh"
#include "fmm_let.cuh"
#include "uj_direct.cuh"
#include "uj_direct_f32.cuh"

using namespace vpm;

struct vpmb200_engine {
    int device = 0;
    int float_bits = 64;
    cudaStream_t stream = nullptr;
    int64_t maxp = 0, ld = 0, np = 0;
    double* state = nullptr;    // NFIELDS x ld
    double* rec = nullptr;      // source records (UJ or E_str pass)
    double* aos = nullptr;      // AoS staging, maxp x NFIELDS
    double* gh_table = nullptr; // Gaussian-erf G/H table
    double* z_table = nullptr;  // Gaussian-erf zeta table
    float* gh_table_f32 = nullptr;
    double* partial = nullptr;  // per-chunk partial outputs of source-split pair launches
    size_t partial_doubles = 0;
    int sm_count = 148;
    int direct_sort = 1;        // Morton-order the direct path internally (vpmb200_set_option)
    int fmm_table_copies = 8;   // UJ_fmm near field: 8 = bank-conflict-free replicated G table, 1 = single copy (vpmb200_set_option)
    int64_t shard_sorted_np = -1;  // >= 0: fmm.perm / sx,sy,sz hold the Morton order of the current local particles
                                   // (set by vpmb200_pack_uj_records, dropped by anything that moves or re-counts them)
    double* probe = nullptr;    // probe scratch: 3 (X) + 3 (U) + 9 (J) rows of probe_ld, + AoS staging
    int64_t probe_cap = 0;
    unsigned long long* counter = nullptr;
    double t = 0.0;
    int64_t nt = 0;
    FmmWorkspace fmm;           // GPU FMM scratch (allocated on first UJ_fmm call)
    FmmLet let;                 // multi-GPU UJ_fmm: local-essential-tree phases (fmm_let.cuh)
    std::vector<int> fmm_lvl;   // cell index range of every tree level
    // DynamicSFS evaluates twice at the SAME positions and strengths (test filter sigma*alpha, then domain filter): the
    // second evaluation reuses the tree, the interaction lists and the local expansions of the first (the far field is
    // the singular kernel when nonzero_sigma = false, so it does not depend on sigma).  do_sfs sets the hint.
    int fmm_hint = 0;           // 1: the next UJ_fmm call should keep its far field; 2: the next call may reuse it
    bool fmm_far_valid = false;
    int64_t fmm_far_np = -1;
    // Static-particle fast path (simulation.jl:355-365): the embedded particles of the current step parked in the state
    // columns [np, np + nstatic) instead of add_particle -> nextstep -> remove_particle; valid while nt == static_gen.
    int64_t nstatic = 0;
    int64_t static_gen = -1;
    int mirror_on = 0;          // method of images (vehicle_vlm_unsteady.jl:245-260): images of every particle join the static set
    double mirror_X[3] = {0, 0, 0}, mirror_n[3] = {0, 0, 1};
    uint64_t launches = 0;      // kernels enqueued by this handle (bench.py's gpu_launches)
    vpmb200_schemes sch;
    std::string err;
};

namespace {

thread_local std::string g_create_error;

int32_t fail(vpmb200_engine* e, int32_t code, const std::string& msg) {
    if (e) e->err = msg; else g_create_error = msg;
    return code;
}

#define CU_TRY(e, call)                                                                                   \
    do {                                                                                                  \
        cudaError_t _st = (call);                                                                         \
        if (_st != cudaSuccess)                                                                           \
            return fail((e), VPMB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(_st));         \
    } while (0)

#define CHECK_HANDLE(h) \
    if (!(h)) return VPMB200_EINVAL

inline int64_t round_up(int64_t x, int64_t m) { return (x + m - 1) / m * m; }
inline unsigned blocks_for(int64_t n, int bt) { return (unsigned)((n + bt - 1) / bt); }

int32_t ensure_probe(vpmb200_engine* e, int64_t m);
inline int32_t ensure_probe_scratch(vpmb200_engine* e) { return ensure_probe(e, 1024); }

// vpm.FMM(nonzero_sigma): 0 = false (singular far field wherever the acceptance holds), 1 = true with 5 core sizes of clearance
// between the closest points of two cells, k >= 2 = true with k core sizes (g differs from 1 by 3e-2 at 3, 1e-3 at 4, 2e-5 at 5)
double nzs_clearance(const vpmb200_schemes& s) {
    return s.fmm_nonzero_sigma <= 0 ? 0.0 : (s.fmm_nonzero_sigma == 1 ? 5.0 : (double)s.fmm_nonzero_sigma);
}

double zeta0_of(int kernel) {
    switch (kernel) {
    case K_GAUSSIANERF: return CONST1;
    case K_WINCKELMANS: return CONST4 * 7.5;
    case K_GAUSSIAN: return 3 * CONST4;
    default: return 1.0;
    }
}

// Launch geometry of a pair kernel: enough (target block, source chunk) work items for >= ~24 waves of 2 CTAs/SM, so
// the tail of the last wave costs a few percent at most (a 125k-target shard alone is only 1.65 waves).
struct Geometry {
    dim3 grid;
    SplitArgs split;
    int nchunks;
};

cudaError_t make_geometry(vpmb200_engine* e, int64_t nt, int ntiles, int ncomp, Geometry* g, int targets_per_block = UJ_BT) {
    const int nblocks = (int)blocks_for(nt, targets_per_block);
    const int slots = 2 * e->sm_count;
    int nchunks = 1;
    if (nblocks < 24 * slots && ntiles >= 16) {
        nchunks = (24 * slots + nblocks - 1) / nblocks;
        if (nchunks > 64) nchunks = 64;
        if (nchunks > ntiles / 8) nchunks = ntiles / 8;   // keep >= 8 tiles (2048 sources) per work item
        if (nchunks < 1) nchunks = 1;
    }
    g->nchunks = nchunks;
    g->grid = dim3((unsigned)nblocks, (unsigned)nchunks, 1);
    g->split.partial = nullptr;
    g->split.ldp = 0;
    g->split.tiles_per_chunk = ntiles;
    if (nchunks > 1) {
        const int64_t ldp = round_up(nt, 32);
        const size_t need = (size_t)nchunks * ncomp * ldp;
        if (need > e->partial_doubles) {
            if (e->partial) cudaFree(e->partial);
            e->partial = nullptr;
            e->partial_doubles = 0;
            cudaError_t st = cudaMalloc(&e->partial, need * sizeof(double));
            if (st != cudaSuccess) return st;
            e->partial_doubles = need;
        }
        g->split.partial = e->partial;
        g->split.ldp = ldp;
        g->split.tiles_per_chunk = (ntiles + nchunks - 1) / nchunks;
        g->nchunks = (ntiles + g->split.tiles_per_chunk - 1) / g->split.tiles_per_chunk;
        g->grid.y = (unsigned)g->nchunks;
    }
    return cudaSuccess;
}

cudaError_t reduce_split(vpmb200_engine* e, const Geometry& g, int ncomp, int64_t nt, double* dstA, double* dstB, int nfirst,
                         int64_t ldo, int accumulate) {
    if (g.nchunks <= 1 || g.split.partial == nullptr) return cudaSuccess;
    reduce_partials_kernel<<<blocks_for(nt, PK_BT), PK_BT, 0, e->stream>>>(g.split.partial, g.nchunks, ncomp, g.split.ldp, nt,
                                                                         dstA, dstB, nfirst, ldo, accumulate);
    e->launches++;
    return cudaGetLastError();
}

template <int K>
cudaError_t launch_uj(vpmb200_engine* e, const double* rec, int ntiles, const double* tx, const double* ty,
                      const double* tz, int64_t nt, double* U, double* J, int64_t ldo, int accumulate) {
    if (nt <= 0) return cudaSuccess;
    Geometry g;
    cudaError_t st = make_geometry(e, nt, ntiles, 12, &g, e->float_bits == 32 ? F32_TPB : UJ_BT);
    if (st != cudaSuccess) return st;
    if (e->float_bits == 32) {
        auto kfn = uj_direct_f32_kernel<K>;
        size_t smem = uj_f32_smem_bytes(K);
        st = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (st != cudaSuccess) return st;
        kfn<<<g.grid, UJ_BT, smem, e->stream>>>(rec, ntiles, tx, ty, tz, nt, U, J, ldo, accumulate, e->gh_table_f32, g.split);
    } else {
        auto kfn = uj_direct_f64_kernel<K>;
        size_t smem = uj_smem_bytes(K);
        st = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (st != cudaSuccess) return st;
        kfn<<<g.grid, UJ_BT, smem, e->stream>>>(rec, ntiles, tx, ty, tz, nt, U, J, ldo, accumulate, e->gh_table, g.split);
    }
    e->launches++;
    st = cudaGetLastError();
    if (st != cudaSuccess) return st;
    return reduce_split(e, g, 12, nt, U, J, 3, ldo, accumulate);
}

cudaError_t dispatch_uj(vpmb200_engine* e, const double* rec, int ntiles, const double* tx, const double* ty,
                        const double* tz, int64_t nt, double* U, double* J, int64_t ldo, int accumulate) {
    switch (e->sch.kernel) {
    case K_GAUSSIANERF: return launch_uj<K_GAUSSIANERF>(e, rec, ntiles, tx, ty, tz, nt, U, J, ldo, accumulate);
    case K_WINCKELMANS: return launch_uj<K_WINCKELMANS>(e, rec, ntiles, tx, ty, tz, nt, U, J, ldo, accumulate);
    case K_GAUSSIAN: return launch_uj<K_GAUSSIAN>(e, rec, ntiles, tx, ty, tz, nt, U, J, ldo, accumulate);
    default: return launch_uj<K_SINGULAR>(e, rec, ntiles, tx, ty, tz, nt, U, J, ldo, accumulate);
    }
}

template <int K>
cudaError_t launch_estr(vpmb200_engine* e, const double* rec, int ntiles, const double* tx, const double* ty, const double* tz,
                        const double* Jt, int64_t ldj, double* sfs, int64_t ldo, int accumulate, int raw = 0) {
    if (e->np <= 0) return cudaSuccess;
    Geometry g;
    cudaError_t st = make_geometry(e, e->np, ntiles, 3, &g);
    if (st != cudaSuccess) return st;
    auto kfn = estr_direct_f64_kernel<K>;
    size_t smem = estr_smem_bytes(K);
    st = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (st != cudaSuccess) return st;
    kfn<<<g.grid, UJ_BT, smem, e->stream>>>(rec, ntiles, tx, ty, tz, e->np, Jt, ldj, e->sch.transposed, sfs, ldo, e->z_table,
                                          g.split, accumulate, raw);
    e->launches++;
    st = cudaGetLastError();
    if (st != cudaSuccess) return st;
    return reduce_split(e, g, 3, e->np, sfs, sfs, 3, ldo, accumulate);
}

cudaError_t dispatch_estr_at(vpmb200_engine* e, const double* rec, int ntiles, const double* tx, const double* ty,
                             const double* tz, const double* Jt, int64_t ldj, double* sfs, int64_t ldo, int accumulate,
                             int raw = 0) {
    switch (e->sch.kernel) {
    case K_GAUSSIANERF: return launch_estr<K_GAUSSIANERF>(e, rec, ntiles, tx, ty, tz, Jt, ldj, sfs, ldo, accumulate, raw);
    case K_WINCKELMANS: return launch_estr<K_WINCKELMANS>(e, rec, ntiles, tx, ty, tz, Jt, ldj, sfs, ldo, accumulate, raw);
    case K_GAUSSIAN: return launch_estr<K_GAUSSIAN>(e, rec, ntiles, tx, ty, tz, Jt, ldj, sfs, ldo, accumulate, raw);
    default: return launch_estr<K_SINGULAR>(e, rec, ntiles, tx, ty, tz, Jt, ldj, sfs, ldo, accumulate, raw);
    }
}

// out (3 rows, particle order) = sum_q zeta_sigma_q(x_p - x_q) vec_q : the zeta pass (vpm.zeta_direct) on the K2 kernel
int32_t zeta_apply(vpmb200_engine* e, const double* vec, int64_t ldv, double* out, int64_t ldo) {
    if (e->np <= 0) return VPMB200_OK;
    const double* S = e->state;
    const int64_t ld = e->ld;
    pack_zeta_records_kernel<<<blocks_for(e->np, TILE_SRC), TILE_SRC, 0, e->stream>>>(
        e->state, ld, e->np, vec, ldv, zeta0_of(e->sch.kernel), e->sch.kernel == K_GAUSSIANERF ? 1 : 0, e->rec);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    CU_TRY(e, dispatch_estr_at(e, e->rec, (int)vpmb200_tiles_for(e->np), S + (size_t)F_X * ld, S + (size_t)(F_X + 1) * ld,
                               S + (size_t)(F_X + 2) * ld, S + (size_t)F_J * ld, ld, out, ldo, 0, 1));
    return VPMB200_OK;
}

// host-side sum of per-block partials (fixed order)
int32_t dot3(vpmb200_engine* e, const double* a, const double* b, double out[3]) {
    int32_t rc = ensure_probe_scratch(e);
    if (rc) return rc;
    const int nb = 128;
    dot3_partials_kernel<<<nb, 256, 0, e->stream>>>(a, e->ld, b, e->ld, e->state + (size_t)F_STATIC * e->ld, e->np, e->probe);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    double hb[nb * 4];
    CU_TRY(e, cudaMemcpyAsync(hb, e->probe, sizeof(hb), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    out[0] = out[1] = out[2] = 0.0;
    for (int k = 0; k < nb; ++k)
        for (int c = 0; c < 3; ++c) out[c] += hb[k * 4 + c];
    return VPMB200_OK;
}

// CoreSpreading spatial adaptation: see include/vpmb200.h
int32_t do_corespreading_reset(vpmb200_engine* e, int32_t* iters, double* residual3) {
    const vpmb200_schemes& s = e->sch;
    if (iters) *iters = 0;
    if (residual3) residual3[0] = residual3[1] = residual3[2] = 0.0;
    if (e->np <= 0 || !(s.cs_sgm0 > 0)) return VPMB200_OK;
    int32_t rc = ensure_probe_scratch(e);
    if (rc) return rc;
    const int64_t ld = e->ld, n = e->np;
    const unsigned nb = blocks_for(n, PK_BT);
    {   // is any spread core beyond beta sgm0?
        const int nbk = 128;
        sigma_max_partials_kernel<<<nbk, 256, 0, e->stream>>>(e->state, ld, n, e->probe);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        double hb[nbk];
        CU_TRY(e, cudaMemcpyAsync(hb, e->probe, sizeof(hb), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(e, cudaStreamSynchronize(e->stream));
        double m = 0.0;
        for (int k = 0; k < nbk; ++k) m = std::max(m, hb[k]);
        if (!(m / s.cs_sgm0 > s.cs_beta)) return VPMB200_OK;
    }
    double* G = e->state + (size_t)F_GAMMA * ld;
    double* b = e->state + (size_t)F_W * ld;         // target vorticity (also the W output rows)
    double* r = e->state + (size_t)F_M * ld;         // M[:,1]
    double* d = e->state + (size_t)(F_M + 3) * ld;   // M[:,2]
    double* Ad = e->state + (size_t)(F_M + 6) * ld;  // M[:,3]
    const double* stat = e->state + (size_t)F_STATIC * ld;
    if ((rc = zeta_apply(e, G, ld, b, ld))) return rc;                       // omega with the spread cores
    set_sigma_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, ld, n, s.cs_sgm0);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    if ((rc = zeta_apply(e, G, ld, Ad, ld))) return rc;                      // A Gamma with the reset cores
    cg_init_kernel<<<nb, PK_BT, 0, e->stream>>>(b, Ad, r, d, ld, stat, n);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    double rr[3];
    if ((rc = dot3(e, r, r, rr))) return rc;
    int32_t it = 0;
    for (; it < s.cs_itmax; ++it) {
        if (std::sqrt(rr[0]) < s.cs_tol && std::sqrt(rr[1]) < s.cs_tol && std::sqrt(rr[2]) < s.cs_tol) break;
        if ((rc = zeta_apply(e, d, ld, Ad, ld))) return rc;
        double dAd[3], rrn[3];
        if ((rc = dot3(e, d, Ad, dAd))) return rc;
        Vec3 alpha, beta;
        for (int k = 0; k < 3; ++k) alpha.v[k] = dAd[k] != 0 ? rr[k] / dAd[k] : 0.0;
        cg_update_kernel<<<nb, PK_BT, 0, e->stream>>>(G, r, d, Ad, ld, stat, n, alpha);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        if ((rc = dot3(e, r, r, rrn))) return rc;
        for (int k = 0; k < 3; ++k) beta.v[k] = rr[k] != 0 ? rrn[k] / rr[k] : 0.0;
        cg_direction_kernel<<<nb, PK_BT, 0, e->stream>>>(d, r, ld, stat, n, beta);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        for (int k = 0; k < 3; ++k) rr[k] = rrn[k];
    }
    if (iters) *iters = it;
    if (residual3)
        for (int k = 0; k < 3; ++k) residual3[k] = std::sqrt(rr[k]);
    return VPMB200_OK;
}

// E_str of the local particles (state rows as targets), accumulated into the state's SFS rows
cudaError_t dispatch_estr(vpmb200_engine* e, const double* rec, int ntiles) {
    const double* S = e->state;
    const int64_t ld = e->ld;
    return dispatch_estr_at(e, rec, ntiles, S + (size_t)F_X * ld, S + (size_t)(F_X + 1) * ld, S + (size_t)(F_X + 2) * ld,
                            S + (size_t)F_J * ld, ld, e->state + (size_t)F_SFS * ld, ld, 1);
}

int32_t zero_rows(vpmb200_engine* e, int first, int count) {
    if (e->np <= 0) return VPMB200_OK;
    zero_rows_kernel<<<blocks_for(e->np, PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, first, count);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

int32_t do_reset(vpmb200_engine* e) {
    int32_t rc = zero_rows(e, F_U, 3);
    if (rc) return rc;
    rc = zero_rows(e, F_J, 9);
    if (rc) return rc;
    return zero_rows(e, F_PSE, 3);
}

int32_t pack_uj(vpmb200_engine* e, double* dst) {
    if (e->np <= 0) return VPMB200_OK;
    pack_uj_records_kernel<<<blocks_for(e->np, TILE_SRC), TILE_SRC, 0, e->stream>>>(e->state, e->ld, e->np, nullptr, dst);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

int32_t pack_estr(vpmb200_engine* e, double* dst) {
    if (e->np <= 0) return VPMB200_OK;
    pack_estr_records_kernel<<<blocks_for(e->np, TILE_SRC), TILE_SRC, 0, e->stream>>>(
        e->state, e->ld, e->np, nullptr, e->sch.transposed, zeta0_of(e->sch.kernel), e->sch.kernel == K_GAUSSIANERF ? 1 : 0, dst);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

int32_t uj_local_from(vpmb200_engine* e, const double* rec, int64_t ntiles_, int accumulate) {
    double* S = e->state;
    const int64_t ld = e->ld;
    int ntiles = (int)ntiles_;
    CU_TRY(e, dispatch_uj(e, rec, ntiles, S + (size_t)F_X * ld, S + (size_t)(F_X + 1) * ld, S + (size_t)(F_X + 2) * ld,
                          e->np, S + (size_t)F_U * ld, S + (size_t)F_J * ld, ld, accumulate));
    return VPMB200_OK;
}

// ---- sharded driver with Morton-ordered LOCAL targets: pack_uj_records sorts the shard once per evaluation; every
//      from_records call then runs the pair kernel on the sorted targets and scatter-adds into the state rows -----------
int32_t shard_sort(vpmb200_engine* e) {
    std::string err;
    if (fmm_reserve_particles(e->fmm, e->np, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    FmmWorkspace& w = e->fmm;
    if (fmm_sort(w, e->state, e->ld, e->np, e->stream, e->launches, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    const unsigned nb = blocks_for(e->np, PK_BT);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X, 1, e->np, w.perm, w.sx, w.lds);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X + 1, 1, e->np, w.perm, w.sy, w.lds);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X + 2, 1, e->np, w.perm, w.sz, w.lds);
    CU_TRY(e, cudaGetLastError());
    e->launches += 3;
    e->shard_sorted_np = e->np;
    return VPMB200_OK;
}

bool shard_is_sorted(const vpmb200_engine* e) {
    return e->direct_sort && e->shard_sorted_np == e->np && e->np >= 4 * TILE_SRC && e->np < 2000000000LL;
}

int32_t uj_local_from_sorted(vpmb200_engine* e, const double* rec, int64_t ntiles, int accumulate) {
    FmmWorkspace& w = e->fmm;
    const unsigned nb = blocks_for(e->np, PK_BT);
    CU_TRY(e, dispatch_uj(e, rec, (int)ntiles, w.sx, w.sy, w.sz, e->np, w.sU, w.sJ, w.lds, 0));
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sU, w.lds, 3, e->np, w.perm, e->state + (size_t)F_U * e->ld, e->ld, accumulate);
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sJ, w.lds, 9, e->np, w.perm, e->state + (size_t)F_J * e->ld, e->ld, accumulate);
    CU_TRY(e, cudaGetLastError());
    e->launches += 2;
    return VPMB200_OK;
}

int32_t estr_local_from_sorted(vpmb200_engine* e, const double* rec, int64_t ntiles) {
    FmmWorkspace& w = e->fmm;
    const unsigned nb = blocks_for(e->np, PK_BT);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_J, 9, e->np, w.perm, w.sJ, w.lds);   // current total J
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    CU_TRY(e, dispatch_estr_at(e, rec, (int)ntiles, w.sx, w.sy, w.sz, w.sJ, w.lds, w.sE, w.lds, 0));
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sE, w.lds, 3, e->np, w.perm, e->state + (size_t)F_SFS * e->ld, e->ld, 1);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

// pfield.UJ(pfield; ...) through the GPU FMM (fmm.cuh): U, J [and the near-field E_str] of every particle
int32_t do_uj_fmm(vpmb200_engine* e, int reset, int reset_sfs, int sfs) {
    e->shard_sorted_np = -1;
    e->fmm.halo = FmmHalo();   // (a local-essential-tree evaluation may have left halo buffers attached)
    e->let.halo_mode = false;
    const int hint = e->fmm_hint;
    const bool far_was_valid = e->fmm_far_valid;
    e->fmm_hint = 0;
    e->fmm_far_valid = false;
    const vpmb200_schemes& s = e->sch;
    if (s.fmm_p < 2 || s.fmm_p > 6) return fail(e, VPMB200_ENOTSUP, "FMM expansion order p must be in 2..6");
    if (s.fmm_ncrit < 1 || s.fmm_ncrit > FMM_MAX_NCRIT) return fail(e, VPMB200_EINVAL, "FMM ncrit must be in 1..256");
    if (!(s.fmm_theta > 0.0 && s.fmm_theta < 1.0)) return fail(e, VPMB200_EINVAL, "FMM theta must be in (0, 1)");
    if (e->float_bits != 64) return fail(e, VPMB200_ENOTSUP, "UJ_fmm runs in FP64 only");
    if (e->np > 2000000000LL) return fail(e, VPMB200_ECAPACITY, "UJ_fmm indexes particles with 32-bit integers");
    int32_t rc;
    if (reset && (rc = zero_rows(e, F_PSE, 3))) return rc;
    if (reset_sfs && (rc = zero_rows(e, F_SFS, 3))) return rc;
    if (e->np <= 0) return VPMB200_OK;
    std::string err;
    const bool reuse = hint == 2 && far_was_valid && !s.fmm_nonzero_sigma && e->fmm_far_np == e->np;
    FmmWorkspace& w = e->fmm;
    if (reuse) {
        CU_TRY(e, fmm_regather(w, e->state, e->ld, e->np, e->stream, e->launches));
    } else {
        if (fmm_reserve(e->fmm, e->np, s.fmm_ncrit, FmmOps<6>::NM, FmmOps<6>::NL, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
        cudaError_t st = fmm_build(w, e->state, e->ld, e->np, s.fmm_ncrit, s.fmm_theta, nzs_clearance(s), e->fmm_lvl, e->stream, e->launches, err);
        if (st != cudaSuccess) return fail(e, st == cudaErrorMemoryAllocation ? VPMB200_ECAPACITY : VPMB200_ECUDA, err);
    }
    const int block = e->fmm_table_copies;   // geometry selector of the near-field kernels (fmm_host.cuh: leaves_uj)
    CU_TRY(e, fmm_evaluate(w, s.fmm_p, s.kernel, block, e->gh_table, e->fmm_lvl, e->stream, e->launches, reuse));
    if (hint == 1 && !s.fmm_nonzero_sigma) {
        e->fmm_far_valid = true;
        e->fmm_far_np = e->np;
    }
    const unsigned nb = blocks_for(e->np, PK_BT);
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sU, w.lds, 3, e->np, w.perm, e->state + (size_t)F_U * e->ld, e->ld, reset ? 0 : 1);
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sJ, w.lds, 9, e->np, w.perm, e->state + (size_t)F_J * e->ld, e->ld, reset ? 0 : 1);
    CU_TRY(e, cudaGetLastError());
    e->launches += 2;
    if (sfs) {
        fmm_gather_estr_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, w.perm, w.sJ, w.lds, s.transposed,
                                                          zeta0_of(s.kernel), w.rec);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        CU_TRY(e, fmm_estr(w, s.kernel, block, s.transposed, e->z_table, e->stream, e->launches));
        fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sE, w.lds, 3, e->np, w.perm, e->state + (size_t)F_SFS * e->ld, e->ld, 1);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
    }
    return VPMB200_OK;
}

// Direct path with internal Morton ordering: targets and source tiles are both visited in Morton order, so a CTA's 256
// targets and a tile's 256 sources are compact boxes and the tile-level far/near classification of K1/K2 stays
// effective for ANY input order (random fields, freshly shed particles...).  Results are scattered back to particle order.
int32_t do_uj_direct_sorted(vpmb200_engine* e, int reset, int sfs) {
    e->shard_sorted_np = -1;   // the shared sort scratch is about to be reused
    std::string err;
    if (fmm_reserve(e->fmm, e->np, 50, FmmOps<6>::NM, FmmOps<6>::NL, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    FmmWorkspace& w = e->fmm;
    if (fmm_sort(w, e->state, e->ld, e->np, e->stream, e->launches, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    const int64_t n = e->np;
    const unsigned nb = blocks_for(n, PK_BT);
    const int ntiles = (int)vpmb200_tiles_for(n);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X, 1, n, w.perm, w.sx, w.lds);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X + 1, 1, n, w.perm, w.sy, w.lds);
    gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X + 2, 1, n, w.perm, w.sz, w.lds);
    pack_uj_records_kernel<<<blocks_for(n, TILE_SRC), TILE_SRC, 0, e->stream>>>(e->state, e->ld, n, w.perm, e->rec);
    CU_TRY(e, cudaGetLastError());
    e->launches += 4;
    CU_TRY(e, dispatch_uj(e, e->rec, ntiles, w.sx, w.sy, w.sz, n, w.sU, w.sJ, w.lds, 0));
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sU, w.lds, 3, n, w.perm, e->state + (size_t)F_U * e->ld, e->ld, reset ? 0 : 1);
    fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sJ, w.lds, 9, n, w.perm, e->state + (size_t)F_J * e->ld, e->ld, reset ? 0 : 1);
    CU_TRY(e, cudaGetLastError());
    e->launches += 2;
    if (sfs) {
        // the E_str pass needs the field's CURRENT (total) J at targets and sources
        gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_J, 9, n, w.perm, w.sJ, w.lds);
        pack_estr_records_kernel<<<blocks_for(n, TILE_SRC), TILE_SRC, 0, e->stream>>>(
            e->state, e->ld, n, w.perm, e->sch.transposed, zeta0_of(e->sch.kernel), e->sch.kernel == K_GAUSSIANERF ? 1 : 0, e->rec);
        CU_TRY(e, cudaGetLastError());
        e->launches += 2;
        CU_TRY(e, dispatch_estr_at(e, e->rec, ntiles, w.sx, w.sy, w.sz, w.sJ, w.lds, w.sE, w.lds, 0));
        fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sE, w.lds, 3, n, w.perm, e->state + (size_t)F_SFS * e->ld, e->ld, 1);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
    }
    return VPMB200_OK;
}

int32_t check_fmm_settings(vpmb200_engine* e) {
    const vpmb200_schemes& s = e->sch;
    if (s.fmm_p < 2 || s.fmm_p > 6) return fail(e, VPMB200_ENOTSUP, "FMM expansion order p must be in 2..6");
    if (s.fmm_ncrit < 1 || s.fmm_ncrit > FMM_MAX_NCRIT) return fail(e, VPMB200_EINVAL, "FMM ncrit must be in 1..256");
    if (!(s.fmm_theta > 0.0 && s.fmm_theta < 1.0)) return fail(e, VPMB200_EINVAL, "FMM theta must be in (0, 1)");
    if (e->float_bits != 64) return fail(e, VPMB200_ENOTSUP, "UJ_fmm runs in FP64 only");
    return VPMB200_OK;
}

// Multi-GPU FMM building block: G is a 24-row mini-state (rows as in the particle record: X 0:3, Gamma 3:6, sigma 6, U 9:12,
// J 15:24; rows 12:15 double as E_str scratch) holding ALL ntot particles of the job.  Every rank builds the same tree
// (deterministic) and evaluates only its share of the leaves; rows outside its share are written as zeros so the ranks'
// results combine with one all-reduce.  pass 0: tree + U, J.   pass 1: near-field E_str from the (reduced) J rows.
int32_t do_fmm_global(vpmb200_engine* e, double* G, int64_t ldg, int64_t ntot, int part, int nparts, int pass) {
    e->shard_sorted_np = -1;
    e->fmm.halo = FmmHalo();
    e->let.halo_mode = false;
    const vpmb200_schemes& s = e->sch;
    int32_t rc = check_fmm_settings(e);
    if (rc) return rc;
    if (ntot <= 0) return VPMB200_OK;
    if (ntot > 2000000000LL) return fail(e, VPMB200_ECAPACITY, "UJ_fmm indexes particles with 32-bit integers");
    std::string err;
    if (fmm_reserve(e->fmm, ntot, s.fmm_ncrit, FmmOps<6>::NM, FmmOps<6>::NL, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    FmmWorkspace& w = e->fmm;
    const int block = e->fmm_table_copies;
    const unsigned nb = blocks_for(ntot, PK_BT);
    if (pass == 0) {
        cudaError_t st = fmm_build(w, G, ldg, ntot, s.fmm_ncrit, s.fmm_theta, nzs_clearance(s), e->fmm_lvl, e->stream,
                                   e->launches, err, part, nparts);
        if (st != cudaSuccess) return fail(e, st == cudaErrorMemoryAllocation ? VPMB200_ECAPACITY : VPMB200_ECUDA, err);
        CU_TRY(e, cudaMemsetAsync(w.sU, 0, sizeof(double) * 3 * w.lds, e->stream));
        CU_TRY(e, cudaMemsetAsync(w.sJ, 0, sizeof(double) * 9 * w.lds, e->stream));
        CU_TRY(e, fmm_evaluate(w, s.fmm_p, s.kernel, block, e->gh_table, e->fmm_lvl, e->stream, e->launches));
        fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sU, w.lds, 3, ntot, w.perm, G + (size_t)F_U * ldg, ldg, 0);
        fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sJ, w.lds, 9, ntot, w.perm, G + (size_t)F_J * ldg, ldg, 0);
        CU_TRY(e, cudaGetLastError());
        e->launches += 2;
    } else {
        if (w.ncells <= 0 || w.nleaves <= 0) return fail(e, VPMB200_EINVAL, "fmm_global pass 1 needs pass 0 first");
        fmm_gather_estr_kernel<<<nb, PK_BT, 0, e->stream>>>(G, ldg, ntot, w.perm, w.sJ, w.lds, s.transposed, zeta0_of(s.kernel), w.rec);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        CU_TRY(e, cudaMemsetAsync(w.sE, 0, sizeof(double) * 3 * w.lds, e->stream));
        CU_TRY(e, fmm_estr(w, s.kernel, block, s.transposed, e->z_table, e->stream, e->launches));
        fmm_scatter_kernel<<<nb, PK_BT, 0, e->stream>>>(w.sE, w.lds, 3, ntot, w.perm, G + (size_t)F_W * ldg, ldg, 0);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
    }
    return VPMB200_OK;
}

// pfield.UJ(pfield; reset, reset_sfs, sfs)
int32_t do_uj(vpmb200_engine* e, int reset, int reset_sfs, int sfs) {
    if (e->sch.uj == VPMB200_UJ_FMM) return do_uj_fmm(e, reset, reset_sfs, sfs);
    int32_t rc;
    if (reset && (rc = zero_rows(e, F_PSE, 3))) return rc;
    if (reset_sfs && (rc = zero_rows(e, F_SFS, 3))) return rc;
    if (e->np <= 0) return VPMB200_OK;
    if (e->direct_sort && e->np >= 4 * TILE_SRC && e->np < 2000000000LL) return do_uj_direct_sorted(e, reset, sfs);
    if ((rc = pack_uj(e, e->rec))) return rc;
    if ((rc = uj_local_from(e, e->rec, vpmb200_tiles_for(e->np), reset ? 0 : 1))) return rc;
    if (sfs) {
        if ((rc = pack_estr(e, e->rec))) return rc;
        CU_TRY(e, dispatch_estr(e, e->rec, (int)vpmb200_tiles_for(e->np)));
    }
    return VPMB200_OK;
}

int32_t do_stage(vpmb200_engine* e, int stage, double a, double b, double dt, const double* Uinf, int relax_inline) {
    if (stage == VPMB200_STAGE_UPDATE) e->shard_sorted_np = -1;   // positions move
    if (e->np <= 0) return VPMB200_OK;
    const vpmb200_schemes& s = e->sch;
    const unsigned nb = blocks_for(e->np, PK_BT);
    const double zeta0 = zeta0_of(s.kernel);
    switch (stage) {
    case VPMB200_STAGE_SCALE_SIGMA_TEST:
        scale_sigma_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.alpha, 0);
        break;
    case VPMB200_STAGE_STORE_TEST:
        dyn_store_test_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.transposed);
        break;
    case VPMB200_STAGE_SCALE_SIGMA_DOMAIN:
        scale_sigma_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.alpha, 1);
        break;
    case VPMB200_STAGE_DYNAMIC_COEFF:
        dyn_coeff_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.transposed, s.alpha, s.sfs_rlxf, s.minC,
                                                    s.maxC, s.force_positive, zeta0);
        break;
    case VPMB200_STAGE_CONSTANT_COEFF:
        const_coeff_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.Cs);
        break;
    case VPMB200_STAGE_CLIP_CONTROL:
        if (!(s.clippings || s.controls)) return VPMB200_OK;
        if (s.clippings || s.controls)
            clip_control_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.clippings, s.controls, s.f, zeta0,
                                                           e->t, e->nt);
        break;
    case VPMB200_STAGE_ZERO_M:
        zero_m_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np);
        break;
    case VPMB200_STAGE_UPDATE: {
        if (!Uinf) return fail(e, VPMB200_EINVAL, "stage UPDATE needs Uinf");
        UpdateParams p;
        p.a = a; p.b = b; p.dt = dt;
        p.Uinf0 = Uinf[0]; p.Uinf1 = Uinf[1]; p.Uinf2 = Uinf[2];
        p.f = s.f; p.g = s.g; p.zeta0 = zeta0; p.nu = s.nu; p.rlxf = s.rlxf;
        p.transposed = s.transposed; p.viscous = s.viscous; p.relaxation = s.relaxation;
        p.euler = (s.integration == VPMB200_INTEGRATION_EULER);
        p.relax_inline = relax_inline;
        update_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, p);
        break;
    }
    case VPMB200_STAGE_RELAX:
        if (s.relaxation == VPMB200_RELAX_NONE) return VPMB200_OK;
        if (s.relaxation != VPMB200_RELAX_NONE)
            relax_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, s.relaxation, s.rlxf);
        break;
    default:
        return fail(e, VPMB200_EINVAL, "unknown stage id");
    }
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

// pfield.SFS(pfield; a, b)   
int32_t do_sfs(vpmb200_engine* e, double a, double b) {
    (void)b;
    const vpmb200_schemes& s = e->sch;
    const bool first = (a == 1.0 || a == 0.0);
    int32_t rc;
    switch (s.sfs) {
    case VPMB200_SFS_NONE:
        return do_uj(e, 1, 0, 0);
    case VPMB200_SFS_CONSTANT:
        if ((rc = do_uj(e, 1, 1, 1))) return rc;
        if (first) {
            if ((rc = do_stage(e, VPMB200_STAGE_CONSTANT_COEFF, 0, 0, 0, nullptr, 0))) return rc;
            if ((rc = do_stage(e, VPMB200_STAGE_CLIP_CONTROL, 0, 0, 0, nullptr, 0))) return rc;
        }
        return VPMB200_OK;
    case VPMB200_SFS_DYNAMIC:
        if (!first) return do_uj(e, 1, 1, 1);
        if ((rc = do_stage(e, VPMB200_STAGE_SCALE_SIGMA_TEST, 0, 0, 0, nullptr, 0))) return rc;
        e->fmm_hint = 1;   // UJ_fmm: keep the tree and the far field ...
        if ((rc = do_uj(e, 1, 1, 1))) return rc;
        if ((rc = do_stage(e, VPMB200_STAGE_STORE_TEST, 0, 0, 0, nullptr, 0))) return rc;
        if ((rc = do_stage(e, VPMB200_STAGE_SCALE_SIGMA_DOMAIN, 0, 0, 0, nullptr, 0))) return rc;
        e->fmm_hint = 2;   // ... only sigma changed in between: same positions, strengths, lists and local expansions
        rc = do_uj(e, 1, 1, 1);
        e->fmm_hint = 0;
        e->fmm_far_valid = false;
        if (rc) return rc;
        if ((rc = do_stage(e, VPMB200_STAGE_DYNAMIC_COEFF, 0, 0, 0, nullptr, 0))) return rc;
        return do_stage(e, VPMB200_STAGE_CLIP_CONTROL, 0, 0, 0, nullptr, 0);
    default:
        return fail(e, VPMB200_EINVAL, "unknown SFS scheme id");
    }
}

int32_t check_schemes(vpmb200_engine* e, const vpmb200_schemes* s) {
    if (s->kernel < 0 || s->kernel > 3) return fail(e, VPMB200_EINVAL, "kernel id out of range");
    if (s->relaxation < 0 || s->relaxation > 2) return fail(e, VPMB200_EINVAL, "relaxation id out of range");
    if (s->sfs < 0 || s->sfs > 2) return fail(e, VPMB200_EINVAL, "SFS id out of range");
    if (s->viscous < 0 || s->viscous > 1) return fail(e, VPMB200_EINVAL, "viscous id out of range");
    if (s->integration < 0 || s->integration > 1) return fail(e, VPMB200_EINVAL, "integration id out of range");
    if (s->uj < 0 || s->uj > 1) return fail(e, VPMB200_EINVAL, "UJ id out of range");
    if (s->sfs == VPMB200_SFS_DYNAMIC && (s->minC < 0 || s->maxC < s->minC))
        return fail(e, VPMB200_EINVAL, "DynamicSFS needs 0 <= minC <= maxC");
    if (s->sfs == VPMB200_SFS_DYNAMIC && !(s->alpha > 0)) return fail(e, VPMB200_EINVAL, "DynamicSFS needs alpha > 0");
    if (s->controls & ~(VPMB200_CTRL_DIRECTIONAL | VPMB200_CTRL_MAGNITUDE))
        return fail(e, VPMB200_ENOTSUP, "control_sigmasensor is not implemented (upstream form unverified)");
    if (s->viscous == VPMB200_VISCOUS_CORESPREADING && s->cs_sgm0 > 0 && (s->cs_itmax < 0 || !(s->cs_beta > 0) || !(s->cs_tol >= 0)))
        return fail(e, VPMB200_EINVAL, "CoreSpreading needs beta > 0, itmax >= 0, tol >= 0");
    if (s->viscous == VPMB200_VISCOUS_CORESPREADING && s->kernel != VPMB200_KERNEL_GAUSSIANERF)
        return fail(e, VPMB200_EINVAL, "CoreSpreading requires the gaussianerf kernel (vpm._kernel_compatibility)");
    return VPMB200_OK;
}

int32_t ensure_probe(vpmb200_engine* e, int64_t m) {
    if (m <= e->probe_cap) return VPMB200_OK;
    if (e->probe) cudaFree(e->probe);
    e->probe = nullptr;
    e->probe_cap = 0;
    int64_t cap = round_up(m, 1024);
    // rows: 3 X + 3 U + 9 J (SoA, ld = cap) followed by AoS staging of 12 * cap
    CU_TRY(e, cudaMalloc(&e->probe, sizeof(double) * (size_t)cap * (15 + 12)));
    e->probe_cap = cap;
    return VPMB200_OK;
}

// AoS [n][nc] <-> SoA rows of length ld (small helper kernels for probes)
__global__ void split_rows_kernel(const double* __restrict__ aos, int nc, int64_t n, double* __restrict__ soa, int64_t ld) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < nc; ++c) soa[(size_t)c * ld + i] = aos[i * nc + c];
}
__global__ void join_rows_kernel(const double* __restrict__ soa, int64_t ld, int nc, int64_t n, double* __restrict__ aos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int c = 0; c < nc; ++c) aos[i * nc + c] = soa[(size_t)c * ld + i];
}

// ---- static-particle fast path + method of images -----------------------------------------------------------------------
__global__ void set_static_flag_kernel(double* __restrict__ soa, int64_t ld, int64_t i0, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) soa[(size_t)F_STATIC * ld + i0 + i] = 1.0;
}

// Image of particle i (i < n) in the plane (X0, nrm) -> column n + i, exactly as the reference writes it
// (src/FLOWUnsteady_vehicle_vlm_unsteady.jl:248-258, src/FLOWUnsteady_simulation.jl:520-533):
//   Xm = X - dot(2 (X - X0), nrm) nrm,   Gm = 2 dot(G, nrm) G / |G| - G   (sic),   sigma, vol, circulation, C copied, static.
__global__ void mirror_images_kernel(double* __restrict__ soa, int64_t ld, int64_t n, Vec3 X0, Vec3 nrm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t o = n + i;
    for (int f = 0; f < NFIELDS; ++f) soa[(size_t)f * ld + o] = 0.0;
    double x[3], g[3];
    for (int c = 0; c < 3; ++c) { x[c] = soa[(size_t)(F_X + c) * ld + i]; g[c] = soa[(size_t)(F_GAMMA + c) * ld + i]; }
    // no contraction: the oracle (and the reference's Julia) round every product and sum
    double a[3];
    for (int c = 0; c < 3; ++c) a[c] = __dmul_rn(2.0, __dsub_rn(x[c], X0.v[c]));
    const double d = __dadd_rn(__dadd_rn(__dmul_rn(a[0], nrm.v[0]), __dmul_rn(a[1], nrm.v[1])), __dmul_rn(a[2], nrm.v[2]));
    const double gn = __dadd_rn(__dadd_rn(__dmul_rn(g[0], nrm.v[0]), __dmul_rn(g[1], nrm.v[1])), __dmul_rn(g[2], nrm.v[2]));
    const double gnorm = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(g[0], g[0]), __dmul_rn(g[1], g[1])), __dmul_rn(g[2], g[2])));
    for (int c = 0; c < 3; ++c) {
        soa[(size_t)(F_X + c) * ld + o] = __dsub_rn(x[c], __dmul_rn(d, nrm.v[c]));
        soa[(size_t)(F_GAMMA + c) * ld + o] = __dsub_rn(__ddiv_rn(__dmul_rn(__dmul_rn(2.0, gn), g[c]), gnorm), g[c]);
        soa[(size_t)(F_C + c) * ld + o] = soa[(size_t)(F_C + c) * ld + i];
    }
    soa[(size_t)F_SIGMA * ld + o] = soa[(size_t)F_SIGMA * ld + i];
    soa[(size_t)F_VOL * ld + o] = soa[(size_t)F_VOL * ld + i];
    soa[(size_t)F_CIRC * ld + o] = soa[(size_t)F_CIRC * ld + i];
    soa[(size_t)F_STATIC * ld + o] = 1.0;
}

// sigma of particle 0 times (or over) `f`, k times in sequence — the reference's Vvpm_on_Xs quirk (simulation.jl:507-513, 555-561)
__global__ void scale_sigma0_kernel(double* __restrict__ soa, int64_t ld, double f, int64_t k, int divide) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double s = soa[(size_t)F_SIGMA * ld];
    for (int64_t j = 0; j < k; ++j) s = divide ? __ddiv_rn(s, f) : __dmul_rn(s, f);
    soa[(size_t)F_SIGMA * ld] = s;
}

__global__ void count_static_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, unsigned long long* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool st = i < n && soa[(size_t)F_STATIC * ld + i] > 0;
    const unsigned m = __ballot_sync(0xffffffffu, st);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(out, (unsigned long long)__popc(m));
}

inline void drop_statics(vpmb200_engine* e) {
    e->nstatic = 0;
    e->static_gen = -1;
}

// images of the first n columns -> columns [n, 2n)
int32_t append_images(vpmb200_engine* e, int64_t n) {
    if (n <= 0) return VPMB200_OK;
    if (2 * n > e->maxp) return fail(e, VPMB200_ECAPACITY, "mirroring needs room for one image per particle (max_particles >= 2 x particles)");
    Vec3 X0, nr;
    for (int c = 0; c < 3; ++c) { X0.v[c] = e->mirror_X[c]; nr.v[c] = e->mirror_n[c]; }
    mirror_images_kernel<<<blocks_for(n, PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, n, X0, nr);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

// The parked static set (while it belongs to the current step) and, with mirroring on, the images of every particle become
// part of the field for the duration of one call; statics_end restores the particle count (and consumes the set).
int32_t statics_begin(vpmb200_engine* e, int64_t* np_saved) {
    *np_saved = e->np;
    if (e->nstatic > 0 && e->static_gen != e->nt) drop_statics(e);   // a stale set is never used
    int64_t n = e->np + e->nstatic;
    if (e->mirror_on && n > 0) {
        int32_t rc = append_images(e, n);
        if (rc) return rc;
        n *= 2;
    }
    if (n != e->np) {
        e->np = n;
        e->shard_sorted_np = -1;
    }
    return VPMB200_OK;
}

void statics_end(vpmb200_engine* e, int64_t np_saved, bool consume) {
    if (e->np != np_saved) e->shard_sorted_np = -1;
    e->np = np_saved;
    if (consume) drop_statics(e);
}

// Instrumentation (not on the hot path): how many (target block, source tile) pairs K1 sends to its branch-free far loop —
// and K2 skips — for the CURRENT field.  One CTA per target block recomputes the box exactly as the pair kernels do.
__global__ void __launch_bounds__(UJ_BT) tile_class_count_kernel(const double* __restrict__ srec, int ntiles,
                                                                 const double* __restrict__ tx, const double* __restrict__ ty,
                                                                 const double* __restrict__ tz, int64_t nt,
                                                                 unsigned long long* __restrict__ far_count) {
    __shared__ PairSmem sm;
    __shared__ int block_far;
    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * UJ_BT + tid;
    const bool live = i < nt;
    if (tid == 0) block_far = 0;
    cta_target_box(sm, live, live ? tx[i] : 0.0, live ? ty[i] : 0.0, live ? tz[i] : 0.0);
    int mine = 0;
    for (int k = tid; k < ntiles; k += UJ_BT) {
        const double* hdr = srec + (size_t)k * TILE_DOUBLES + TILE_HDR;
        double h[7];
        for (int c = 0; c < 7; ++c) h[c] = hdr[c];
        mine += box_dist2(sm.tbox, h) > h[6] ? 1 : 0;
    }
    atomicAdd(&block_far, mine);
    __syncthreads();
    if (tid == 0) atomicAdd(far_count, (unsigned long long)block_far);
}

}  // namespace

extern "C" {

const char* vpmb200_version(void) { return "vpmb200 0.1 (sm_100a)"; }

int32_t vpmb200_default_schemes(vpmb200_schemes* s) {
    if (!s) return VPMB200_EINVAL;
    std::memset(s, 0, sizeof(*s));
    s->kernel = VPMB200_KERNEL_GAUSSIANERF;
    s->f = 0.0;
    s->g = 1.0 / 5.0;
    s->transposed = 1;
    s->relaxation = VPMB200_RELAX_PEDRIZZETTI;
    s->rlxf = 0.3;
    s->sfs = VPMB200_SFS_NONE;
    s->alpha = 0.999;
    s->sfs_rlxf = 0.005;
    s->minC = 0.0;
    s->maxC = 1.0;
    s->Cs = 1.0;
    s->viscous = VPMB200_VISCOUS_INVISCID;
    s->integration = VPMB200_INTEGRATION_RK3;
    s->cs_sgm0 = 0.0;
    s->cs_beta = 1.5;
    s->cs_itmax = 15;
    s->cs_tol = 1e-3;
    s->uj = VPMB200_UJ_DIRECT;
    s->fmm_p = 4;
    s->fmm_ncrit = 50;
    s->fmm_theta = 0.4;
    s->fmm_nonzero_sigma = 0;
    return VPMB200_OK;
}

int32_t vpmb200_create(int64_t max_particles, int32_t nfields, int32_t float_bits, int32_t device, vpmb200_handle* out) {
    if (!out) return VPMB200_EINVAL;
    *out = nullptr;
    if (max_particles <= 0) return fail(nullptr, VPMB200_EINVAL, "max_particles must be positive");
    if (nfields != NFIELDS) return fail(nullptr, VPMB200_EINVAL, "nfields must be 43");
    if (float_bits != 64 && float_bits != 32) return fail(nullptr, VPMB200_EINVAL, "float_bits must be 64 or 32");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0)
        return fail(nullptr, VPMB200_ENODEVICE, "no CUDA device: the engine has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(nullptr, VPMB200_EINVAL, "device ordinal out of range");
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
        return fail(nullptr, VPMB200_ENODEVICE, "device is not sm_100 (Blackwell B200); kernels are built for sm_100a only");
    vpmb200_engine* e = new (std::nothrow) vpmb200_engine();
    if (!e) return VPMB200_EINVAL;
    e->device = device;
    e->sm_count = prop.multiProcessorCount;
    e->float_bits = float_bits;
    e->maxp = max_particles;
    e->ld = round_up(max_particles, 256);
    vpmb200_default_schemes(&e->sch);
#define CREATE_TRY(call)                                                                         \
    do {                                                                                         \
        cudaError_t _st = (call);                                                                \
        if (_st != cudaSuccess) {                                                                \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(_st);                \
            vpmb200_destroy(e);                                                                  \
            return VPMB200_ECUDA;                                                                \
        }                                                                                        \
    } while (0)
    CREATE_TRY(cudaSetDevice(device));
    CREATE_TRY(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
    CREATE_TRY(cudaMalloc(&e->state, sizeof(double) * (size_t)NFIELDS * e->ld));
    CREATE_TRY(cudaMemsetAsync(e->state, 0, sizeof(double) * (size_t)NFIELDS * e->ld, e->stream));
    CREATE_TRY(cudaMalloc(&e->rec, sizeof(double) * (size_t)(vpmb200_tiles_for(e->maxp) * TILE_DOUBLES)));
    CREATE_TRY(cudaMalloc(&e->aos, sizeof(double) * (size_t)NFIELDS * e->maxp));
    CREATE_TRY(cudaMalloc(&e->gh_table, sizeof(vpm_gt_GH)));
    CREATE_TRY(cudaMalloc(&e->z_table, sizeof(vpm_gt_Z)));
    CREATE_TRY(cudaMalloc(&e->gh_table_f32, sizeof(vpm_gt32_GH)));
    CREATE_TRY(cudaMalloc(&e->counter, sizeof(unsigned long long)));
    CREATE_TRY(cudaMemcpyAsync(e->gh_table, vpm_gt_GH, sizeof(vpm_gt_GH), cudaMemcpyHostToDevice, e->stream));
    CREATE_TRY(cudaMemcpyAsync(e->z_table, vpm_gt_Z, sizeof(vpm_gt_Z), cudaMemcpyHostToDevice, e->stream));
    CREATE_TRY(cudaMemcpyAsync(e->gh_table_f32, vpm_gt32_GH, sizeof(vpm_gt32_GH), cudaMemcpyHostToDevice, e->stream));
    CREATE_TRY(cudaStreamSynchronize(e->stream));
#undef CREATE_TRY
    *out = e;
    return VPMB200_OK;
}

int32_t vpmb200_destroy(vpmb200_handle e) {
    CHECK_HANDLE(e);
    cudaSetDevice(e->device);
    if (e->stream) cudaStreamSynchronize(e->stream);
    cudaFree(e->state);
    cudaFree(e->rec);
    cudaFree(e->aos);
    cudaFree(e->gh_table);
    cudaFree(e->z_table);
    cudaFree(e->gh_table_f32);
    cudaFree(e->probe);
    cudaFree(e->partial);
    cudaFree(e->counter);
    fmm_free(e->fmm);
    let_free(e->let);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
    return VPMB200_OK;
}

const char* vpmb200_last_error(vpmb200_handle e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int32_t vpmb200_set_schemes(vpmb200_handle e, const vpmb200_schemes* s) {
    CHECK_HANDLE(e);
    if (!s) return fail(e, VPMB200_EINVAL, "schemes is NULL");
    int32_t rc = check_schemes(e, s);
    if (rc) return rc;
    e->sch = *s;
    return VPMB200_OK;
}

int32_t vpmb200_get_schemes(vpmb200_handle e, vpmb200_schemes* s) {
    CHECK_HANDLE(e);
    if (!s) return fail(e, VPMB200_EINVAL, "schemes is NULL");
    *s = e->sch;
    return VPMB200_OK;
}

int32_t vpmb200_set_time(vpmb200_handle e, double t, int64_t nt) {
    CHECK_HANDLE(e);
    e->t = t;
    e->nt = nt;
    return VPMB200_OK;
}

int32_t vpmb200_get_time(vpmb200_handle e, double* t, int64_t* nt) {
    CHECK_HANDLE(e);
    if (t) *t = e->t;
    if (nt) *nt = e->nt;
    return VPMB200_OK;
}

int32_t vpmb200_get_np(vpmb200_handle e, int64_t* np) {
    CHECK_HANDLE(e);
    if (!np) return fail(e, VPMB200_EINVAL, "np is NULL");
    *np = e->np;
    return VPMB200_OK;
}

// Contiguous runs [c0, c1) of particle columns selected by a field-group mask (group table of include/vpmb200.h).
static int mask_runs(uint32_t mask, int runs[13][2]) {
    static const int first[14] = {F_X, F_GAMMA, F_SIGMA, F_VOL, F_CIRC, F_U, F_W, F_J, F_PSE, F_M, F_C, F_SFS, F_STATIC, NFIELDS};
    int n = 0;
    for (int g = 0; g < 13; ++g) {
        if (!((mask >> g) & 1u)) continue;
        if (n > 0 && runs[n - 1][1] == first[g]) runs[n - 1][1] = first[g + 1];
        else { runs[n][0] = first[g]; runs[n][1] = first[g + 1]; ++n; }
    }
    return n;
}

static int32_t upload_block(vpmb200_engine* e, const double* particles, int64_t ld, int64_t n, int64_t dst0, uint32_t mask) {
    e->shard_sorted_np = -1;
    drop_statics(e);   // the columns behind np are about to change hands
    if (n <= 0) return VPMB200_OK;
    if (ld < NFIELDS) return fail(e, VPMB200_EINVAL, "ld < 43");
    CU_TRY(e, cudaSetDevice(e->device));
    // host (ld per particle) -> device AoS staging (NFIELDS per particle): only the selected column runs cross the bus
    int runs[13][2];
    const int nr = mask_runs(mask & VPMB200_FM_ALL, runs);
    for (int r = 0; r < nr; ++r)
        CU_TRY(e, cudaMemcpy2DAsync(e->aos + runs[r][0], sizeof(double) * NFIELDS, particles + runs[r][0], sizeof(double) * ld,
                                    sizeof(double) * (runs[r][1] - runs[r][0]), (size_t)n, cudaMemcpyHostToDevice, e->stream));
    dim3 grid(blocks_for(n, 32), (NFIELDS + 31) / 32), block(32, 8);
    aos_to_soa_kernel<<<grid, block, 0, e->stream>>>(e->aos, NFIELDS, n, e->state, e->ld, dst0, mask);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    // the borrowed host pointer must not be read after return
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    return VPMB200_OK;
}

int32_t vpmb200_upload(vpmb200_handle e, const double* particles, int64_t ld, int64_t np, uint32_t field_mask) {
    CHECK_HANDLE(e);
    if (np < 0) return fail(e, VPMB200_EINVAL, "np < 0");
    if (np > e->maxp) return fail(e, VPMB200_ECAPACITY, "np exceeds max_particles");
    if (np > 0 && !particles) return fail(e, VPMB200_EINVAL, "particles is NULL");
    int32_t rc = upload_block(e, particles, ld, np, 0, field_mask);
    if (rc) return rc;
    e->np = np;
    return VPMB200_OK;
}

int32_t vpmb200_download(vpmb200_handle e, double* particles, int64_t ld, int64_t np, uint32_t field_mask) {
    CHECK_HANDLE(e);
    if (np < 0 || np > e->np) return fail(e, VPMB200_EINVAL, "np out of range");
    if (np == 0) return VPMB200_OK;
    if (!particles) return fail(e, VPMB200_EINVAL, "particles is NULL");
    if (ld < NFIELDS) return fail(e, VPMB200_EINVAL, "ld < 43");
    CU_TRY(e, cudaSetDevice(e->device));
    dim3 grid(blocks_for(np, 32), (NFIELDS + 31) / 32), block(32, 8);
    soa_to_aos_kernel<<<grid, block, 0, e->stream>>>(e->state, e->ld, np, e->aos, NFIELDS, field_mask);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    // columns not selected keep the host's values: only the selected column runs are written (strided 2-D copies)
    int runs[13][2];
    const int nr = mask_runs(field_mask & VPMB200_FM_ALL, runs);
    for (int r = 0; r < nr; ++r)
        CU_TRY(e, cudaMemcpy2DAsync(particles + runs[r][0], sizeof(double) * ld, e->aos + runs[r][0], sizeof(double) * NFIELDS,
                                    sizeof(double) * (runs[r][1] - runs[r][0]), (size_t)np, cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    return VPMB200_OK;
}

int32_t vpmb200_host_register(void* ptr, uint64_t bytes) {
    if (!ptr || bytes == 0) return fail(nullptr, VPMB200_EINVAL, "host_register: NULL pointer or zero size");
    cudaError_t st = cudaHostRegister(ptr, (size_t)bytes, cudaHostRegisterPortable);
    if (st == cudaErrorHostMemoryAlreadyRegistered) {
        cudaGetLastError();
        return VPMB200_OK;
    }
    if (st == cudaErrorNoDevice || st == cudaErrorInsufficientDriver) {
        cudaGetLastError();
        return fail(nullptr, VPMB200_ENODEVICE, "host_register: no CUDA device");
    }
    if (st != cudaSuccess) {
        cudaGetLastError();
        return fail(nullptr, VPMB200_ECUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(st));
    }
    return VPMB200_OK;
}

int32_t vpmb200_host_unregister(void* ptr) {
    if (!ptr) return fail(nullptr, VPMB200_EINVAL, "host_unregister: NULL pointer");
    cudaError_t st = cudaHostUnregister(ptr);
    if (st != cudaSuccess && st != cudaErrorHostMemoryNotRegistered) {
        cudaGetLastError();
        return fail(nullptr, VPMB200_ECUDA, std::string("cudaHostUnregister: ") + cudaGetErrorString(st));
    }
    cudaGetLastError();
    return VPMB200_OK;
}

int32_t vpmb200_add_particles(vpmb200_handle e, const double* cols, int64_t ld, int64_t n) {
    CHECK_HANDLE(e);
    if (n < 0) return fail(e, VPMB200_EINVAL, "n < 0");
    if (e->np + n > e->maxp) return fail(e, VPMB200_ECAPACITY, "adding particles would exceed max_particles");
    if (n > 0 && !cols) return fail(e, VPMB200_EINVAL, "cols is NULL");
    int32_t rc = upload_block(e, cols, ld, n, e->np, VPMB200_FM_ALL);
    if (rc) return rc;
    e->np += n;
    return VPMB200_OK;
}

int32_t vpmb200_remove_particle(vpmb200_handle e, int64_t i) {
    CHECK_HANDLE(e);
    if (i < 0 || i >= e->np) return fail(e, VPMB200_EINVAL, "particle index out of range");
    e->shard_sorted_np = -1;
    drop_statics(e);
    CU_TRY(e, cudaSetDevice(e->device));
    if (i != e->np - 1) {
        move_column_kernel<<<1, 64, 0, e->stream>>>(e->state, e->ld, i, e->np - 1);
        CU_TRY(e, cudaGetLastError());
    e->launches++;
    }
    e->np -= 1;
    return VPMB200_OK;
}

int32_t vpmb200_remove_where(vpmb200_handle e, int32_t criterion, const double* params, int64_t* removed) {
    CHECK_HANDLE(e);
    if (!params) return fail(e, VPMB200_EINVAL, "params is NULL");
    static const int nparams[5] = {0, 2, 2, 9, 4};
    if (criterion < 1 || criterion > 4) return fail(e, VPMB200_EINVAL, "unknown removal criterion");
    e->shard_sorted_np = -1;
    drop_statics(e);
    if (removed) *removed = 0;
    const int64_t n = e->np;
    if (n <= 0) return VPMB200_OK;
    if (n > 2000000000LL) return fail(e, VPMB200_ECAPACITY, "remove_where indexes particles with 32-bit integers");
    CU_TRY(e, cudaSetDevice(e->device));
    std::string err;
    if (fmm_reserve(e->fmm, n, 50, FmmOps<6>::NM, FmmOps<6>::NL, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    FmmWorkspace& w = e->fmm;   // scratch: perm = keep flags, perm_alt = exclusive scan, count_at = removed slots,
                                //          leaf_flag-sized buffers are too small, so f / g live in keys / keys_alt
    RemoveCriterion c;
    c.kind = criterion;
    for (int k = 0; k < 9; ++k) c.p[k] = k < nparams[criterion] ? params[k] : 0.0;
    const unsigned nb = blocks_for(n, PK_BT);
    keep_flags_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, n, c, w.perm);
    CU_TRY(e, cudaGetLastError());
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, w.perm, w.perm_alt, (int)n, e->stream);
    if (fmm_cub(w, tb, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
    tb = w.cub_bytes;
    CU_TRY(e, cub::DeviceScan::ExclusiveSum(w.cub_tmp, tb, w.perm, w.perm_alt, (int)n, e->stream));
    int last_scan = 0, last_flag = 0;
    CU_TRY(e, cudaMemcpyAsync(&last_scan, w.perm_alt + n - 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaMemcpyAsync(&last_flag, w.perm + n - 1, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    e->launches += 2;
    const int64_t K = (int64_t)last_scan + last_flag;
    if (K < n) {
        if (K > 0) {
            int* f = reinterpret_cast<int*>(w.keys);
            int* g = reinterpret_cast<int*>(w.keys_alt);
            int* changed = reinterpret_cast<int*>(w.counters);
            list_removed_kernel<<<nb, PK_BT, 0, e->stream>>>(w.perm, w.perm_alt, n, w.count_at);
            wrap_map_kernel<<<nb, PK_BT, 0, e->stream>>>(w.perm_alt, n, K, w.count_at, f);
            e->launches += 2;
            for (int it = 0; it < 40; ++it) {   // pointer doubling: at most log2(n) + 1 rounds
                int h = 0;
                CU_TRY(e, cudaMemcpyAsync(changed, &h, sizeof(int), cudaMemcpyHostToDevice, e->stream));
                wrap_double_kernel<<<nb, PK_BT, 0, e->stream>>>(f, n, g, changed);
                e->launches++;
                CU_TRY(e, cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, e->stream));
                CU_TRY(e, cudaStreamSynchronize(e->stream));
                std::swap(f, g);
                if (!h) break;
            }
            move_survivors_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, w.perm, n, f);
            CU_TRY(e, cudaGetLastError());
            e->launches++;
        }
        e->np = K;
    }
    if (removed) *removed = n - K;
    return VPMB200_OK;
}

int32_t vpmb200_zeta(vpmb200_handle e) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    return zeta_apply(e, e->state + (size_t)F_GAMMA * e->ld, e->ld, e->state + (size_t)F_W * e->ld, e->ld);
}

int32_t vpmb200_corespreading_reset(vpmb200_handle e, int32_t* iters, double* residual3) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    return do_corespreading_reset(e, iters, residual3);
}

int32_t vpmb200_monitors(vpmb200_handle e, double* out6) {
    CHECK_HANDLE(e);
    if (!out6) return fail(e, VPMB200_EINVAL, "out is NULL");
    for (int k = 0; k < 6; ++k) out6[k] = 0.0;
    if (e->np <= 0) return VPMB200_OK;
    CU_TRY(e, cudaSetDevice(e->device));
    int32_t rc = ensure_probe(e, 1024);
    if (rc) return rc;
    const int nb = 128;
    monitor_partials_kernel<<<nb, 256, 0, e->stream>>>(e->state, e->ld, e->np, e->probe);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    double hb[nb * 8];
    CU_TRY(e, cudaMemcpyAsync(hb, e->probe, sizeof(hb), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    double acc[6] = {0, 0, 0, 0, 0, 0};
    for (int b = 0; b < nb; ++b)
        for (int k = 0; k < 6; ++k) acc[k] += hb[b * 8 + k];
    out6[0] = acc[0];
    const double cnt = acc[3];
    out6[1] = cnt > 0 ? acc[1] / cnt : 0.0;
    out6[2] = cnt > 1 ? std::sqrt(std::max(0.0, (acc[2] - acc[1] * acc[1] / cnt) / (cnt - 1))) : 0.0;
    out6[3] = cnt;
    out6[4] = acc[4];
    out6[5] = acc[5];
    return VPMB200_OK;
}

int32_t vpmb200_reset_particles(vpmb200_handle e) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    return do_reset(e);
}

int32_t vpmb200_reset_particles_sfs(vpmb200_handle e) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    return zero_rows(e, F_SFS, 3);
}

int32_t vpmb200_uj(vpmb200_handle e, int32_t reset, int32_t reset_sfs, int32_t sfs) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    int64_t np0;
    int32_t rc = statics_begin(e, &np0);
    if (rc == VPMB200_OK) rc = do_uj(e, reset, reset_sfs, sfs);
    statics_end(e, np0, false);
    return rc;
}

static int32_t probe_eval(vpmb200_engine* e, const double* X, int64_t m, double* U, double* J) {
    // probes always use the direct kernel: m targets x np sources
    int32_t rc = ensure_probe(e, m);
    if (rc) return rc;
    const int64_t pl = e->probe_cap;
    double* soaX = e->probe;
    double* soaU = e->probe + 3 * pl;
    double* soaJ = e->probe + 6 * pl;
    double* stage = e->probe + 15 * pl;
    CU_TRY(e, cudaMemcpyAsync(stage, X, sizeof(double) * 3 * (size_t)m, cudaMemcpyHostToDevice, e->stream));
    split_rows_kernel<<<blocks_for(m, PK_BT), PK_BT, 0, e->stream>>>(stage, 3, m, soaX, pl);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    if ((rc = pack_uj(e, e->rec))) return rc;
    int ntiles = (int)vpmb200_tiles_for(e->np);
    CU_TRY(e, dispatch_uj(e, e->rec, ntiles, soaX, soaX + pl, soaX + 2 * pl, m, soaU, soaJ, pl, 0));
    join_rows_kernel<<<blocks_for(m, PK_BT), PK_BT, 0, e->stream>>>(soaU, pl, 3, m, stage);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    CU_TRY(e, cudaMemcpyAsync(U, stage, sizeof(double) * 3 * (size_t)m, cudaMemcpyDeviceToHost, e->stream));
    if (J) {
        double* stageJ = stage + 3 * pl;
        join_rows_kernel<<<blocks_for(m, PK_BT), PK_BT, 0, e->stream>>>(soaJ, pl, 9, m, stageJ);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        CU_TRY(e, cudaMemcpyAsync(J, stageJ, sizeof(double) * 9 * (size_t)m, cudaMemcpyDeviceToHost, e->stream));
    }
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    return VPMB200_OK;
}

int32_t vpmb200_uj_probe(vpmb200_handle e, const double* X, int64_t m, double* U, double* J) {
    return vpmb200_uj_probe_ex(e, X, m, 1.0, 0, U, J);
}

int32_t vpmb200_uj_probe_ex(vpmb200_handle e, const double* X, int64_t m, double fsgm, int32_t mirror, double* U, double* J) {
    CHECK_HANDLE(e);
    if (m < 0) return fail(e, VPMB200_EINVAL, "m < 0");
    if (m == 0) return VPMB200_OK;
    if (!X || !U) return fail(e, VPMB200_EINVAL, "X or U is NULL");
    if (!(fsgm > 0)) return fail(e, VPMB200_EINVAL, "fsgm must be positive (simulation.jl:211-213)");
    CU_TRY(e, cudaSetDevice(e->device));
    // Vvpm_on_Xs's "singularize" loop (simulation.jl:507-513) indexes the particle MATRIX with one subscript, so instead of
    // scaling every static particle's core it multiplies sigma of particle 1 by fsgm once per static particle present in the
    // field at that moment (normally none: the statics are added after the loop).  Reproduced literally, undone after.
    int64_t k = 0;
    if (std::fabs(fsgm) != 1.0 && e->np > 0) {
        CU_TRY(e, cudaMemsetAsync(e->counter, 0, sizeof(unsigned long long), e->stream));
        count_static_kernel<<<blocks_for(round_up(e->np, 32), PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, e->counter);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        unsigned long long c = 0;
        CU_TRY(e, cudaMemcpyAsync(&c, e->counter, sizeof(c), cudaMemcpyDeviceToHost, e->stream));
        CU_TRY(e, cudaStreamSynchronize(e->stream));
        k = (int64_t)c;
        if (k > 0) {
            scale_sigma0_kernel<<<1, 32, 0, e->stream>>>(e->state, e->ld, fsgm, k, 0);
            CU_TRY(e, cudaGetLastError());
            e->launches++;
        }
    }
    int64_t np0;
    int32_t rc = statics_begin(e, &np0);          // static_particles_fun (+ its images)
    if (rc == VPMB200_OK && mirror && e->np > 0) {   // Vvpm_on_Xs's own method of images on top (simulation.jl:518-535)
        rc = append_images(e, e->np);
        if (rc == VPMB200_OK) e->np *= 2;
    }
    if (rc == VPMB200_OK) rc = probe_eval(e, X, m, U, J);
    statics_end(e, np0, false);
    if (k > 0) {
        scale_sigma0_kernel<<<1, 32, 0, e->stream>>>(e->state, e->ld, fsgm, k, 1);
        cudaError_t st = cudaGetLastError();
        if (st != cudaSuccess && rc == VPMB200_OK) rc = fail(e, VPMB200_ECUDA, cudaGetErrorString(st));
        e->launches++;
    }
    return rc;
}

int32_t vpmb200_sfs(vpmb200_handle e, double a, double b) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    return do_sfs(e, a, b);
}

static int32_t nextstep_body(vpmb200_engine* e, double dt, const double* Uinf, int32_t relax) {
    int32_t rc;
    if (e->np > 0) {
        if (e->sch.integration == VPMB200_INTEGRATION_EULER) {
            if ((rc = do_sfs(e, 1.0, 1.0))) return rc;
            if ((rc = do_stage(e, VPMB200_STAGE_UPDATE, 0.0, 1.0, dt, Uinf, relax ? 1 : 0))) return rc;
            if (e->sch.viscous == VPMB200_VISCOUS_CORESPREADING && (rc = do_corespreading_reset(e, nullptr, nullptr))) return rc;
        } else {
            static const double AB[3][2] = {{0.0, 1.0 / 3.0}, {-5.0 / 9.0, 15.0 / 16.0}, {-153.0 / 128.0, 8.0 / 15.0}};
            if ((rc = do_stage(e, VPMB200_STAGE_ZERO_M, 0, 0, 0, nullptr, 0))) return rc;
            for (int st = 0; st < 3; ++st) {
                if ((rc = do_sfs(e, AB[st][0], AB[st][1]))) return rc;
                if ((rc = do_stage(e, VPMB200_STAGE_UPDATE, AB[st][0], AB[st][1], dt, Uinf, 0))) return rc;
            }
            // spatial adaptation is checked once the last substep is done
            if (e->sch.viscous == VPMB200_VISCOUS_CORESPREADING && (rc = do_corespreading_reset(e, nullptr, nullptr))) return rc;
            if (relax && e->sch.relaxation != VPMB200_RELAX_NONE) {
                if ((rc = do_uj(e, 1, 0, 0))) return rc;
                if ((rc = do_stage(e, VPMB200_STAGE_RELAX, 0, 0, 0, nullptr, 0))) return rc;
            }
        }
    }
    return VPMB200_OK;
}

int32_t vpmb200_nextstep(vpmb200_handle e, double dt, const double* Uinf, int32_t relax) {
    CHECK_HANDLE(e);
    if (!Uinf) return fail(e, VPMB200_EINVAL, "Uinf is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    // the reference's loop: static_particles_function -> nextstep -> remove the statics (simulation.jl:355-365); here the
    // parked set (vpmb200_set_statics) joins the field for the step and is dropped with it
    int64_t np0;
    int32_t rc = statics_begin(e, &np0);
    if (rc == VPMB200_OK) rc = nextstep_body(e, dt, Uinf, relax);
    statics_end(e, np0, true);
    if (rc) return rc;
    e->t += dt;
    e->nt += 1;
    return VPMB200_OK;
}

int32_t vpmb200_set_statics(vpmb200_handle e, const double* cols, int64_t ld, int64_t n, int64_t generation) {
    CHECK_HANDLE(e);
    if (n < 0 || (n > 0 && !cols)) return fail(e, VPMB200_EINVAL, "set_statics: bad arguments");
    if (e->np + n > e->maxp) return fail(e, VPMB200_ECAPACITY, "static particles would exceed max_particles");
    CU_TRY(e, cudaSetDevice(e->device));
    int32_t rc = upload_block(e, cols, ld, n, e->np, VPMB200_FM_ALL);   // (drops the previous set)
    if (rc) return rc;
    if (n > 0) {
        set_static_flag_kernel<<<blocks_for(n, PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, n);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
    }
    e->nstatic = n;
    e->static_gen = generation;
    return VPMB200_OK;
}

int32_t vpmb200_get_statics(vpmb200_handle e, int64_t* n, int64_t* generation) {
    CHECK_HANDLE(e);
    if (n) *n = e->nstatic;
    if (generation) *generation = e->static_gen;
    return VPMB200_OK;
}

int32_t vpmb200_set_mirror(vpmb200_handle e, int32_t enabled, const double* X0, const double* normal) {
    CHECK_HANDLE(e);
    if (enabled && (!X0 || !normal)) return fail(e, VPMB200_EINVAL, "set_mirror: plane point or normal is NULL");
    e->mirror_on = enabled != 0;
    if (X0 && normal)   // the plane is also what vpmb200_uj_probe_ex(mirror = 1) uses
        for (int c = 0; c < 3; ++c) { e->mirror_X[c] = X0[c]; e->mirror_n[c] = normal[c]; }
    return VPMB200_OK;
}

int32_t vpmb200_count_nonfinite(vpmb200_handle e, int64_t* count) {
    CHECK_HANDLE(e);
    if (!count) return fail(e, VPMB200_EINVAL, "count is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    CU_TRY(e, cudaMemsetAsync(e->counter, 0, sizeof(unsigned long long), e->stream));
    if (e->np > 0) {
        count_nonfinite_kernel<<<blocks_for(e->np, PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, e->counter);
        CU_TRY(e, cudaGetLastError());
    e->launches++;
    }
    unsigned long long c = 0;
    CU_TRY(e, cudaMemcpyAsync(&c, e->counter, sizeof(c), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    *count = (int64_t)c;
    return VPMB200_OK;
}

int32_t vpmb200_device_field(vpmb200_handle e, int32_t field, double** ptr, int64_t* ld) {
    CHECK_HANDLE(e);
    if (field < 0 || field >= NFIELDS || !ptr) return fail(e, VPMB200_EINVAL, "bad field index");
    *ptr = e->state + (size_t)field * e->ld;
    if (ld) *ld = e->ld;
    return VPMB200_OK;
}

int32_t vpmb200_stream(vpmb200_handle e, void** stream) {
    CHECK_HANDLE(e);
    if (!stream) return fail(e, VPMB200_EINVAL, "stream is NULL");
    *stream = (void*)e->stream;
    return VPMB200_OK;
}

int32_t vpmb200_fmm_global(vpmb200_handle e, double* G, int64_t ldg, int64_t ntot, int32_t part, int32_t nparts, int32_t pass) {
    CHECK_HANDLE(e);
    if (!G || ldg < ntot || nparts < 1 || part < 0 || part >= nparts || pass < 0 || pass > 1)
        return fail(e, VPMB200_EINVAL, "bad fmm_global arguments");
    CU_TRY(e, cudaSetDevice(e->device));
    return do_fmm_global(e, G, ldg, ntot, part, nparts, pass);
}

// ---- multi-GPU UJ_fmm: local-essential-tree phases (fmm_let.cuh; the collectives between them are the caller's) ----------
#define LET_TRY(e, call)                                                                         \
    do {                                                                                         \
        std::string _err;                                                                        \
        auto _f = [&](std::string& err) -> cudaError_t { return (call); };                       \
        cudaError_t _st = _f(_err);                                                              \
        if (_st != cudaSuccess)                                                                  \
            return fail((e), _st == cudaErrorMemoryAllocation ? VPMB200_ECAPACITY : VPMB200_ECUDA, \
                        _err.empty() ? std::string(cudaGetErrorString(_st)) : _err);             \
    } while (0)

int32_t vpmb200_let_cell_bytes(void) { return (int32_t)sizeof(FmmCell); }

int32_t vpmb200_let_bounds(vpmb200_handle e, double* lohi6) {
    CHECK_HANDLE(e);
    if (!lohi6) return fail(e, VPMB200_EINVAL, "lohi is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    int32_t rc = check_fmm_settings(e);
    if (rc) return rc;
    if (e->np > 2000000000LL) return fail(e, VPMB200_ECAPACITY, "UJ_fmm indexes particles with 32-bit integers");
    e->shard_sorted_np = -1;
    LET_TRY(e, let_bounds(e->fmm, e->state, e->ld, e->np, lohi6, e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_keys(vpmb200_handle e, const double* lohi6_global, int32_t Lc, void** hist_dev, void** binmax_dev) {
    CHECK_HANDLE(e);
    if (!lohi6_global || Lc < 1 || Lc > LET_MAX_LC) return fail(e, VPMB200_EINVAL, "let_keys: bad arguments (1 <= Lc <= 6)");
    CU_TRY(e, cudaSetDevice(e->device));
    FmmRoot cube;
    if (!let_cube_from_bounds(lohi6_global, &cube)) return fail(e, VPMB200_EINVAL, "particle positions are not finite");
    LET_TRY(e, let_keys(e->fmm, e->let, e->state, e->ld, e->np, cube, Lc, e->sch.fmm_nonzero_sigma != 0, e->stream, e->launches, err));
    if (hist_dev) *hist_dev = e->let.hist;
    if (binmax_dev) *binmax_dev = e->sch.fmm_nonzero_sigma ? (void*)e->let.binmax : nullptr;
    return VPMB200_OK;
}

int32_t vpmb200_let_partition(vpmb200_handle e, int32_t nparts, int32_t part, int32_t use_work, int64_t* send_counts) {
    CHECK_HANDLE(e);
    if (nparts < 1 || part < 0 || part >= nparts || !send_counts) return fail(e, VPMB200_EINVAL, "let_partition: bad arguments");
    if (!e->let.hist) return fail(e, VPMB200_EINVAL, "let_partition needs let_keys first");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_partition(e->let, nparts, part, e->sch.fmm_ncrit, use_work != 0, send_counts, e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_cut(const int32_t* hist, const int64_t* work, int32_t Lc, int32_t ncrit, int32_t nparts, uint64_t* splitters) {
    if (!hist || !splitters || Lc < 1 || Lc > LET_MAX_LC || ncrit < 1 || nparts < 1 || nparts > 1024) return VPMB200_EINVAL;
    static_assert(sizeof(long long) == sizeof(int64_t), "work counts are 64-bit");
    std::vector<uint64_t> sp;
    let_cut(hist, reinterpret_cast<const long long*>(work), Lc, ncrit, nparts, sp);
    for (int k = 0; k <= nparts; ++k) splitters[k] = sp[k];
    return VPMB200_OK;
}

int32_t vpmb200_let_work(vpmb200_handle e, void** work_dev) {
    CHECK_HANDLE(e);
    if (!work_dev) return fail(e, VPMB200_EINVAL, "work_dev is NULL");
    *work_dev = e->let.work;
    return VPMB200_OK;
}

int32_t vpmb200_let_pack(vpmb200_handle e, double* rows) {
    CHECK_HANDLE(e);
    if (e->np > 0 && !rows) return fail(e, VPMB200_EINVAL, "rows is NULL");
    if (e->let.n_home != e->np) return fail(e, VPMB200_EINVAL, "let_pack needs let_keys of the current particles first");
    CU_TRY(e, cudaSetDevice(e->device));
    if (e->np > 0) {
        let_pack_rows_kernel<<<blocks_for(e->np, PK_BT), PK_BT, 0, e->stream>>>(e->state, e->ld, e->np, e->let.hperm, rows);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
    }
    return VPMB200_OK;
}

int32_t vpmb200_let_build(vpmb200_handle e, const double* rows, int64_t n_own, int64_t n_all, int32_t reuse, int64_t* info4) {
    CHECK_HANDLE(e);
    if (n_own < 0 || n_all < n_own || (n_own > 0 && !rows)) return fail(e, VPMB200_EINVAL, "let_build: bad arguments");
    if (n_all > 2000000000LL) return fail(e, VPMB200_ECAPACITY, "UJ_fmm indexes particles with 32-bit integers");
    CU_TRY(e, cudaSetDevice(e->device));
    int32_t rc = check_fmm_settings(e);
    if (rc) return rc;
    const vpmb200_schemes& s = e->sch;
    LET_TRY(e, let_build(e->fmm, e->let, rows, n_own, n_all, s.fmm_ncrit, nzs_clearance(s), s.fmm_p, reuse != 0,
                         e->stream, e->launches, err));
    if (info4) {
        info4[0] = e->let.ncells_own;
        info4[1] = e->fmm.nleaves;
        info4[2] = 3 * let_nm(s.fmm_p);
        info4[3] = e->let.n_own;
    }
    return VPMB200_OK;
}

int32_t vpmb200_let_ptrs(vpmb200_handle e, void** ptrs3) {
    CHECK_HANDLE(e);
    if (!ptrs3) return fail(e, VPMB200_EINVAL, "ptrs is NULL");
    ptrs3[0] = e->fmm.cells;
    ptrs3[1] = e->fmm.M;
    ptrs3[2] = e->fmm.rec;
    return VPMB200_OK;
}

int32_t vpmb200_let_attach_tree(vpmb200_handle e, const void* cells_recv, const double* M_recv, int64_t slot_cells,
                                const int64_t* ncells, const int64_t* nparticles) {
    CHECK_HANDLE(e);
    if (!ncells || !nparticles || slot_cells < 0) return fail(e, VPMB200_EINVAL, "let_attach_tree: bad arguments");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_attach_tree(e->fmm, e->let, static_cast<const FmmCell*>(cells_recv), M_recv, slot_cells, ncells, nparticles,
                               e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_attach_skeleton(vpmb200_handle e, const void* cells_recv, int64_t slot_cells, const int64_t* ncells,
                                    const int64_t* nparticles, const int64_t* nleaves) {
    CHECK_HANDLE(e);
    if (!ncells || !nparticles || !nleaves || slot_cells < 0) return fail(e, VPMB200_EINVAL, "let_attach_skeleton: bad arguments");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_attach_skeleton(e->fmm, e->let, static_cast<const FmmCell*>(cells_recv), slot_cells, ncells, nparticles, nleaves,
                                   e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_halo_plan(vpmb200_handle e, int64_t* counts3, void** req_cells_dev, void** req_leaf_dev) {
    CHECK_HANDLE(e);
    if (!counts3) return fail(e, VPMB200_EINVAL, "counts3 is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_halo_plan(e->fmm, e->let, counts3, e->stream, e->launches, err));
    if (req_cells_dev) *req_cells_dev = e->let.req_cells;
    if (req_leaf_dev) *req_leaf_dev = e->let.req_leaf;
    return VPMB200_OK;
}

int32_t vpmb200_let_halo_serve(vpmb200_handle e, const void* req_cells, int64_t ncell, const void* req_leaf, int64_t nleaf,
                               double* M_out, double* rec_out) {
    CHECK_HANDLE(e);
    if (ncell < 0 || nleaf < 0 || (ncell > 0 && M_out && !req_cells) || (nleaf > 0 && rec_out && !req_leaf))
        return fail(e, VPMB200_EINVAL, "let_halo_serve: bad arguments");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_halo_serve(e->fmm, e->let, static_cast<const int*>(req_cells), ncell, static_cast<const int2*>(req_leaf), nleaf,
                              M_out, rec_out, e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_halo_set(vpmb200_handle e, const double* M2, const double* rec2) {
    CHECK_HANDLE(e);
    if (!e->let.halo_mode) return fail(e, VPMB200_EINVAL, "let_halo_set needs let_attach_skeleton first");
    FmmHalo& h = e->fmm.halo;
    h.rec_split = (int)e->let.n_own;
    h.cell_split = e->let.ncells_own;
    h.mslot = e->let.mslot;
    if (M2) h.M2 = M2;
    if (rec2) h.rec2 = rec2;
    return VPMB200_OK;
}

int32_t vpmb200_let_attach_records(vpmb200_handle e, const double* rec_recv, int64_t slot_n, const int64_t* nparticles) {
    CHECK_HANDLE(e);
    if (!nparticles || slot_n < 0) return fail(e, VPMB200_EINVAL, "let_attach_records: bad arguments");
    if ((int)e->let.part_off.size() != e->let.nparts) return fail(e, VPMB200_EINVAL, "let_attach_records needs let_attach_tree first");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_attach_records(e->fmm, e->let, rec_recv, slot_n, nparticles, e->stream, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_evaluate(vpmb200_handle e, double* out_rows, int32_t reuse, int32_t stage) {
    CHECK_HANDLE(e);
    if (!(stage >= 0 && stage <= 2) && stage != 5 && stage != 6)
        return fail(e, VPMB200_EINVAL, "let_evaluate: stage must be 0, 1, 2, 5 or 6");
    if (e->let.n_own > 0 && !out_rows && stage != 1) return fail(e, VPMB200_EINVAL, "out_rows is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    const vpmb200_schemes& s = e->sch;
    if (reuse && (s.fmm_nonzero_sigma || !e->let.far_valid)) return fail(e, VPMB200_EINVAL, "let_evaluate: nothing to reuse");
    if (!reuse && stage != 2 && stage != 6 && e->let.work) CU_TRY(e, cudaMemsetAsync(e->let.work, 0, sizeof(long long) * e->let.bins, e->stream));
    LET_TRY(e, let_evaluate(e->fmm, e->let, s.fmm_theta, nzs_clearance(s), s.kernel, e->fmm_table_copies, e->gh_table,
                            out_rows, reuse != 0, stage, e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_estr_records(vpmb200_handle e) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    CU_TRY(e, let_estr_records(e->fmm, e->let, e->sch.transposed, zeta0_of(e->sch.kernel), e->stream, e->launches));
    return VPMB200_OK;
}

int32_t vpmb200_let_estr_evaluate(vpmb200_handle e, double* out_rows) {
    CHECK_HANDLE(e);
    if (e->let.n_own > 0 && !out_rows) return fail(e, VPMB200_EINVAL, "out_rows is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    LET_TRY(e, let_estr_evaluate(e->fmm, e->let, e->sch.kernel, e->fmm_table_copies, e->sch.transposed, e->z_table, out_rows,
                                 e->stream, e->launches, err));
    return VPMB200_OK;
}

int32_t vpmb200_let_finish(vpmb200_handle e, const double* res_rows, int32_t what, int32_t reset) {
    CHECK_HANDLE(e);
    if (what != 0 && what != 1) return fail(e, VPMB200_EINVAL, "let_finish: what must be 0 (U, J) or 1 (E_str)");
    if (e->np > 0 && !res_rows) return fail(e, VPMB200_EINVAL, "res_rows is NULL");
    if (e->let.n_home != e->np) return fail(e, VPMB200_EINVAL, "let_finish: the particle count changed since let_keys");
    CU_TRY(e, cudaSetDevice(e->device));
    int32_t rc;
    if (what == 0 && reset && (rc = zero_rows(e, F_PSE, 3))) return rc;
    if (e->np <= 0) return VPMB200_OK;
    if (what == 0)
        let_finish_kernel<<<blocks_for(e->np, PK_BT), PK_BT, 0, e->stream>>>(res_rows, 12, e->np, e->let.hperm, e->state, e->ld, F_U, 3,
                                                                           F_J, 9, reset ? 0 : 1);
    else
        let_finish_kernel<<<blocks_for(e->np, PK_BT), PK_BT, 0, e->stream>>>(res_rows, 3, e->np, e->let.hperm, e->state, e->ld, F_SFS, 3,
                                                                           F_SFS, 0, 1);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    return VPMB200_OK;
}

int32_t vpmb200_set_option(vpmb200_handle e, const char* name, int64_t value) {
    CHECK_HANDLE(e);
    if (!name) return fail(e, VPMB200_EINVAL, "option name is NULL");
    if (std::strcmp(name, "direct_sort") == 0) {
        e->direct_sort = value != 0;
        return VPMB200_OK;
    }
    if (std::strcmp(name, "fmm_table_copies") == 0) {
        if (value != 1 && value != 8) return fail(e, VPMB200_EINVAL, "fmm_table_copies must be 1 or 8");
        e->fmm_table_copies = (int)value;
        return VPMB200_OK;
    }
    return fail(e, VPMB200_EINVAL, std::string("unknown option: ") + name);
}

int32_t vpmb200_fmm_stats(vpmb200_handle e, int64_t* stats) {
    CHECK_HANDLE(e);
    if (!stats) return fail(e, VPMB200_EINVAL, "stats is NULL");
    stats[0] = e->fmm.ncells;
    stats[1] = e->fmm.nleaves;
    stats[2] = e->fmm.nlevels;
    stats[3] = e->fmm.n_m2l;
    stats[4] = e->fmm.n_p2p;
    return VPMB200_OK;
}

int32_t vpmb200_direct_tile_stats(vpmb200_handle e, int64_t* stats) {
    CHECK_HANDLE(e);
    if (!stats) return fail(e, VPMB200_EINVAL, "stats is NULL");
    stats[0] = stats[1] = stats[2] = stats[3] = 0;
    if (e->np <= 0) return VPMB200_OK;
    if (e->np >= 2000000000LL) return fail(e, VPMB200_ECAPACITY, "tile statistics index particles with 32-bit integers");
    CU_TRY(e, cudaSetDevice(e->device));
    e->shard_sorted_np = -1;
    const int64_t n = e->np;
    const int ntiles = (int)vpmb200_tiles_for(n);
    const unsigned nb = blocks_for(n, PK_BT);
    const double *tx = e->state + (size_t)F_X * e->ld, *ty = tx + e->ld, *tz = ty + e->ld;
    if (e->direct_sort && n >= 4 * TILE_SRC) {   // the geometry do_uj_direct_sorted uses
        std::string err;
        if (fmm_reserve(e->fmm, n, 50, FmmOps<6>::NM, FmmOps<6>::NL, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
        FmmWorkspace& w = e->fmm;
        if (fmm_sort(w, e->state, e->ld, n, e->stream, e->launches, err) != cudaSuccess) return fail(e, VPMB200_ECUDA, err);
        gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X, 1, n, w.perm, w.sx, w.lds);
        gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X + 1, 1, n, w.perm, w.sy, w.lds);
        gather_rows_kernel<<<nb, PK_BT, 0, e->stream>>>(e->state, e->ld, F_X + 2, 1, n, w.perm, w.sz, w.lds);
        pack_uj_records_kernel<<<blocks_for(n, TILE_SRC), TILE_SRC, 0, e->stream>>>(e->state, e->ld, n, w.perm, e->rec);
        CU_TRY(e, cudaGetLastError());
        e->launches += 4;
        tx = w.sx; ty = w.sy; tz = w.sz;
    } else {
        int32_t rc = pack_uj(e, e->rec);
        if (rc) return rc;
    }
    CU_TRY(e, cudaMemsetAsync(e->counter, 0, sizeof(unsigned long long), e->stream));
    const unsigned nblocks = blocks_for(n, UJ_BT);
    tile_class_count_kernel<<<nblocks, UJ_BT, 0, e->stream>>>(e->rec, ntiles, tx, ty, tz, n, e->counter);
    CU_TRY(e, cudaGetLastError());
    e->launches++;
    unsigned long long c = 0;
    CU_TRY(e, cudaMemcpyAsync(&c, e->counter, sizeof(c), cudaMemcpyDeviceToHost, e->stream));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    stats[0] = nblocks;
    stats[1] = ntiles;
    stats[2] = (int64_t)c;
    stats[3] = (int64_t)nblocks * ntiles;
    return VPMB200_OK;
}

int32_t vpmb200_fmm_times(vpmb200_handle e, double* ms6) {
    CHECK_HANDLE(e);
    if (!ms6) return fail(e, VPMB200_EINVAL, "ms is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    for (int k = 0; k < 6; ++k) {
        ms6[k] = 0.0;
        float t = 0.f;
        if (e->fmm.tset[k] && cudaEventElapsedTime(&t, e->fmm.tev[k][0], e->fmm.tev[k][1]) == cudaSuccess) ms6[k] = t;
        else cudaGetLastError();
    }
    return VPMB200_OK;
}

int32_t vpmb200_launch_count(vpmb200_handle e, uint64_t* count) {
    CHECK_HANDLE(e);
    if (!count) return fail(e, VPMB200_EINVAL, "count is NULL");
    *count = e->launches;
    return VPMB200_OK;
}

int32_t vpmb200_synchronize(vpmb200_handle e) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    CU_TRY(e, cudaStreamSynchronize(e->stream));
    return VPMB200_OK;
}

int64_t vpmb200_tiles_for(int64_t nparticles) { return nparticles <= 0 ? 0 : round_up(nparticles, TILE_SRC) / TILE_SRC; }

int64_t vpmb200_tile_doubles(void) { return TILE_DOUBLES; }

int32_t vpmb200_pack_uj_records(vpmb200_handle e, double* dst) {
    CHECK_HANDLE(e);
    if (!dst) return fail(e, VPMB200_EINVAL, "dst is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    if (e->direct_sort && e->np >= 4 * TILE_SRC && e->np < 2000000000LL) {
        // the shard's tiles go out in Morton order (tight tile boxes for every receiver) and its targets are visited in
        // the same order by the from_records calls that follow
        int32_t rc = shard_sort(e);
        if (rc) return rc;
        pack_uj_records_kernel<<<blocks_for(e->np, TILE_SRC), TILE_SRC, 0, e->stream>>>(e->state, e->ld, e->np, e->fmm.perm, dst);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        return VPMB200_OK;
    }
    e->shard_sorted_np = -1;
    return pack_uj(e, dst);
}

int32_t vpmb200_pack_estr_records(vpmb200_handle e, double* dst) {
    CHECK_HANDLE(e);
    if (!dst) return fail(e, VPMB200_EINVAL, "dst is NULL");
    CU_TRY(e, cudaSetDevice(e->device));
    if (shard_is_sorted(e)) {   // positions have not moved since the UJ pass of this evaluation: reuse its ordering
        pack_estr_records_kernel<<<blocks_for(e->np, TILE_SRC), TILE_SRC, 0, e->stream>>>(
            e->state, e->ld, e->np, e->fmm.perm, e->sch.transposed, zeta0_of(e->sch.kernel), e->sch.kernel == K_GAUSSIANERF ? 1 : 0, dst);
        CU_TRY(e, cudaGetLastError());
        e->launches++;
        return VPMB200_OK;
    }
    return pack_estr(e, dst);
}

int32_t vpmb200_uj_from_records(vpmb200_handle e, const double* tiles, int64_t ntiles, int32_t accumulate) {
    CHECK_HANDLE(e);
    if (ntiles < 0 || (ntiles > 0 && !tiles)) return fail(e, VPMB200_EINVAL, "bad tiles");
    CU_TRY(e, cudaSetDevice(e->device));
    if (e->np <= 0) return VPMB200_OK;
    if (shard_is_sorted(e)) return uj_local_from_sorted(e, tiles, ntiles, accumulate);
    return uj_local_from(e, tiles, ntiles, accumulate);  // ntiles == 0 writes zeros unless accumulating
}

int32_t vpmb200_estr_from_records(vpmb200_handle e, const double* tiles, int64_t ntiles) {
    CHECK_HANDLE(e);
    if (ntiles < 0 || (ntiles > 0 && !tiles)) return fail(e, VPMB200_EINVAL, "bad tiles");
    CU_TRY(e, cudaSetDevice(e->device));
    if (e->np <= 0 || ntiles == 0) return VPMB200_OK;
    if (shard_is_sorted(e)) return estr_local_from_sorted(e, tiles, ntiles);
    CU_TRY(e, dispatch_estr(e, tiles, (int)ntiles));
    return VPMB200_OK;
}

int32_t vpmb200_stage(vpmb200_handle e, int32_t stage, double a, double b, double dt, const double* Uinf) {
    CHECK_HANDLE(e);
    CU_TRY(e, cudaSetDevice(e->device));
    if (stage == VPMB200_STAGE_UPDATE_EULER_RELAX) return do_stage(e, VPMB200_STAGE_UPDATE, a, b, dt, Uinf, 1);
    return do_stage(e, stage, a, b, dt, Uinf, 0);
}

}  // extern "C"

#include "multi.inl"
