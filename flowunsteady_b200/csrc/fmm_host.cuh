// fmm_host.cuh — host orchestration of the GPU FMM (fmm.cuh): workspace, tree build loop, traversal loop, passes.
#pragma once

#include <algorithm>
#include <string>
#include <vector>

#include "fmm.cuh"

namespace vpm {

struct FmmWorkspace {
    int64_t cap_n = 0;
    int cap_cells = 0;
    unsigned int cap_pairs = 0, cap_p2p = 0;
    uint64_t *keys = nullptr, *keys_alt = nullptr;
    int *perm = nullptr, *perm_alt = nullptr;
    double *sx = nullptr, *sy = nullptr, *sz = nullptr, *rec = nullptr;  // Morton-ordered targets / source records
    double *sU = nullptr, *sJ = nullptr, *sE = nullptr;                   // Morton-ordered outputs (3, 9, 3 rows of lds)
    int64_t lds = 0;
    FmmCell* cells = nullptr;
    int *nchild = nullptr, *child_off = nullptr, *leaf_flag = nullptr, *leaf_pos = nullptr, *leaves = nullptr;
    double *M = nullptr, *L = nullptr;
    size_t ml_doubles_M = 0, ml_doubles_L = 0;
    uint64_t *front_a = nullptr, *front_b = nullptr, *m2l = nullptr, *m2l_sorted = nullptr, *p2p = nullptr, *p2p_sorted = nullptr;
    unsigned int *m2l_off = nullptr, *p2p_off = nullptr;
    int2* runs = nullptr;          // particle range of every sorted P2P entry
    int* count_at = nullptr;       // leaf particle count, indexed by the leaf's first particle
    int *mine = nullptr, *leaf_keys = nullptr, *leaf_keys_alt = nullptr, *leaves_alt = nullptr;  // multi-GPU ownership
    FmmCounters* counters = nullptr;
    double* bounds = nullptr;
    void* cub_tmp = nullptr;
    size_t cub_bytes = 0;
    // statistics of the last evaluation
    int ncells = 0, nleaves = 0, nlevels = 0;
    int leaf_lo = 0, leaf_hi = 0;   // leaves [leaf_lo, leaf_hi) are evaluated by the near-field / L2P kernels (multi-GPU split)
    unsigned int n_m2l = 0, n_p2p = 0;
    int sms = 0;                    // multiprocessors of the current device (grid of the persistent leaf kernels)
    // local essential tree (fmm_let.cuh): when set, the downward pass and the leaf kernels read cells / multipoles from these
    // combined arrays ([own | other ranks']) instead of w.cells / w.M; targets are always the own cells [0, ncells)
    const FmmCell* cells_eval = nullptr;
    const double* M_eval = nullptr;
    FmmHalo halo;                   // demand-driven LET: received multipoles / records (defaults: off)
    int64_t count_at_cap = 0;       // entries of count_at (the halo mode indexes it beyond the particle count)
    // device time of the sections of the LAST evaluation (CUDA events on the engine's stream; vpmb200_fmm_times):
    // 0 sort + tree, 1 lists (traversal sweeps incl. their host read-backs, sorts), 2 upward, 3 M2L + L2L, 4 L2P + near field,
    // 5 E_str near field
    cudaEvent_t tev[6][2] = {};
    bool tset[6] = {false, false, false, false, false, false};
};

inline void fmm_tic(FmmWorkspace& w, int k, cudaStream_t st) {
    if (!w.tev[k][0]) {
        cudaEventCreate(&w.tev[k][0]);
        cudaEventCreate(&w.tev[k][1]);
    }
    cudaEventRecord(w.tev[k][0], st);
    w.tset[k] = false;
}
inline void fmm_toc(FmmWorkspace& w, int k, cudaStream_t st) {
    if (w.tev[k][1]) {
        cudaEventRecord(w.tev[k][1], st);
        w.tset[k] = true;
    }
}

// Grid of the persistent near-field kernels: as many CTAs as are resident at once (occupancy of `kfn` x SMs), fewer when
// there are not enough leaves; zeroes the leaf counter (counters->next is idle once the traversal is done).
template <typename K>
inline cudaError_t fmm_leaf_grid(FmmWorkspace& w, K kfn, int threads, size_t smem, int ctas_needed, cudaStream_t st, int& grid) {
    int per_sm = 0;
    cudaError_t eo = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kfn, threads, smem);
    if (eo != cudaSuccess) return eo;
    if (w.sms == 0) {
        int dev = 0;
        cudaError_t e = cudaGetDevice(&dev);
        if (e != cudaSuccess) return e;
        if ((e = cudaDeviceGetAttribute(&w.sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
    }
    grid = std::min(ctas_needed, std::max(per_sm, 1) * w.sms);
    return cudaMemsetAsync(&w.counters->next, 0, sizeof(unsigned int), st);
}

inline void fmm_free(FmmWorkspace& w) {
    for (auto& pair : w.tev)
        for (auto& ev : pair)
            if (ev) cudaEventDestroy(ev);
    void* ptrs[] = {w.keys, w.keys_alt, w.perm, w.perm_alt, w.sx, w.sy, w.sz, w.rec, w.sU, w.sJ, w.sE, w.cells, w.nchild,
                    w.child_off, w.leaf_flag, w.leaf_pos, w.leaves, w.M, w.L, w.front_a, w.front_b, w.m2l, w.m2l_sorted,
                    w.p2p, w.p2p_sorted, w.m2l_off, w.p2p_off, w.counters, w.bounds, w.cub_tmp, w.runs, w.count_at, w.mine,
                    w.leaf_keys, w.leaf_keys_alt, w.leaves_alt};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    w = FmmWorkspace();
}

#define FMM_TRY(call)                                        \
    do {                                                     \
        cudaError_t _st = (call);                            \
        if (_st != cudaSuccess) {                            \
            err = std::string(#call) + ": " + cudaGetErrorString(_st); \
            return _st;                                      \
        }                                                    \
    } while (0)

template <typename T>
inline cudaError_t fmm_grow(T*& ptr, size_t& have, size_t need, std::string& err) {
    if (need <= have && ptr) return cudaSuccess;
    if (ptr) cudaFree(ptr);
    ptr = nullptr;
    have = 0;
    // 25 % headroom: the cell count drifts by a fraction of a percent from one evaluation to the next, and an exact fit
    // would pay a cudaFree + cudaMalloc (a device synchronisation each) on every upward drift
    const size_t cap = need + need / 4;
    FMM_TRY(cudaMalloc(&ptr, sizeof(T) * cap));
    have = cap;
    return cudaSuccess;
}

// CUB scratch grows on demand (each CUB call is first asked for its requirement).
inline cudaError_t fmm_cub(FmmWorkspace& w, size_t bytes, std::string& err) {
    if (bytes <= w.cub_bytes && w.cub_tmp) return cudaSuccess;
    if (w.cub_tmp) cudaFree(w.cub_tmp);
    w.cub_tmp = nullptr;
    w.cub_bytes = 0;
    FMM_TRY(cudaMalloc(&w.cub_tmp, bytes + bytes / 4 + 256));   // headroom: list sizes drift between evaluations
    w.cub_bytes = bytes + bytes / 4 + 256;
    return cudaSuccess;
}

// (Re)allocate the pair-list buffers for the current cap_pairs / cap_p2p.
inline cudaError_t fmm_alloc_pairs(FmmWorkspace& w, std::string& err) {
    void** ptrs[] = {(void**)&w.front_a, (void**)&w.front_b, (void**)&w.m2l, (void**)&w.m2l_sorted, (void**)&w.p2p,
                     (void**)&w.p2p_sorted, (void**)&w.runs};
    for (void** p : ptrs) {
        if (*p) cudaFree(*p);
        *p = nullptr;
    }
    FMM_TRY(cudaMalloc(&w.front_a, sizeof(uint64_t) * w.cap_pairs));
    FMM_TRY(cudaMalloc(&w.front_b, sizeof(uint64_t) * w.cap_pairs));
    FMM_TRY(cudaMalloc(&w.m2l, sizeof(uint64_t) * w.cap_pairs));
    FMM_TRY(cudaMalloc(&w.m2l_sorted, sizeof(uint64_t) * w.cap_pairs));
    FMM_TRY(cudaMalloc(&w.p2p, sizeof(uint64_t) * w.cap_p2p));
    FMM_TRY(cudaMalloc(&w.p2p_sorted, sizeof(uint64_t) * w.cap_p2p));
    FMM_TRY(cudaMalloc(&w.runs, sizeof(int2) * w.cap_p2p));
    return cudaSuccess;
}

// Particle-sized scratch (sort keys, Morton-ordered copies).  Used by the FMM, the sorted direct path and remove_where.
inline cudaError_t fmm_reserve_particles(FmmWorkspace& w, int64_t n, std::string& err) {
    if (n <= w.cap_n && w.keys) return cudaSuccess;
    void* ptrs[] = {w.keys, w.keys_alt, w.perm, w.perm_alt, w.sx, w.sy, w.sz, w.rec, w.sU, w.sJ, w.sE, w.count_at};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    w.keys = w.keys_alt = nullptr;
    w.perm = w.perm_alt = w.count_at = nullptr;
    w.sx = w.sy = w.sz = w.rec = w.sU = w.sJ = w.sE = nullptr;
    const int64_t cap_n = std::max<int64_t>(n + n / 8, 4096);
    w.cap_n = 0;
    w.lds = (cap_n + 31) / 32 * 32;
    FMM_TRY(cudaMalloc(&w.keys, sizeof(uint64_t) * cap_n));
    FMM_TRY(cudaMalloc(&w.keys_alt, sizeof(uint64_t) * cap_n));
    FMM_TRY(cudaMalloc(&w.perm, sizeof(int) * cap_n));
    FMM_TRY(cudaMalloc(&w.perm_alt, sizeof(int) * cap_n));
    FMM_TRY(cudaMalloc(&w.sx, sizeof(double) * w.lds));
    FMM_TRY(cudaMalloc(&w.sy, sizeof(double) * w.lds));
    FMM_TRY(cudaMalloc(&w.sz, sizeof(double) * w.lds));
    FMM_TRY(cudaMalloc(&w.rec, sizeof(double) * REC_REALS * cap_n));
    FMM_TRY(cudaMalloc(&w.sU, sizeof(double) * 3 * w.lds));
    FMM_TRY(cudaMalloc(&w.sJ, sizeof(double) * 9 * w.lds));
    FMM_TRY(cudaMalloc(&w.sE, sizeof(double) * 3 * w.lds));
    FMM_TRY(cudaMalloc(&w.count_at, sizeof(int) * cap_n));
    w.count_at_cap = cap_n;
    if (!w.counters) FMM_TRY(cudaMalloc(&w.counters, sizeof(FmmCounters)));
    if (!w.bounds) FMM_TRY(cudaMalloc(&w.bounds, sizeof(double) * 6 * 256));
    w.cap_n = cap_n;
    return cudaSuccess;
}

// Cell-sized arrays; `cells` is a capacity (the tree build restarts with a larger one if it is exceeded).  Grows with 12.5 %
// headroom: the request follows the particle count, and a rank of a sharded field owns a few hundred particles more or fewer
// from one evaluation to the next — an exact fit paid 12 cudaFree + 12 cudaMalloc (7 ms, and once 180 ms, in the tree phase
// of a 2 x 2.5M-particle LET step: profiles/r02x_let_step_probe_2gpu.txt) on every upward drift of the owned count.
inline cudaError_t fmm_reserve_cells(FmmWorkspace& w, int64_t cells, std::string& err) {
    if (cells <= w.cap_cells && w.cells) return cudaSuccess;
    cells += cells / 8 + 1024;
    void* ptrs[] = {w.cells, w.nchild, w.child_off, w.leaf_flag, w.leaf_pos, w.leaves, w.leaves_alt, w.leaf_keys,
                    w.leaf_keys_alt, w.mine, w.m2l_off, w.p2p_off};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    w.cells = nullptr;
    w.nchild = w.child_off = w.leaf_flag = w.leaf_pos = w.leaves = w.leaves_alt = w.leaf_keys = w.leaf_keys_alt = w.mine = nullptr;
    w.m2l_off = w.p2p_off = nullptr;
    w.cap_cells = 0;
    FMM_TRY(cudaMalloc(&w.cells, sizeof(FmmCell) * cells));
    FMM_TRY(cudaMalloc(&w.nchild, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.child_off, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.leaf_flag, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.leaf_pos, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.leaves, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.leaves_alt, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.leaf_keys, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.leaf_keys_alt, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.mine, sizeof(int) * cells));
    FMM_TRY(cudaMalloc(&w.m2l_off, sizeof(unsigned int) * (cells + 1)));
    FMM_TRY(cudaMalloc(&w.p2p_off, sizeof(unsigned int) * (cells + 1)));
    w.cap_cells = (int)cells;
    return cudaSuccess;
}

// Back-compat entry used by the sort-only callers: particle scratch only.
inline cudaError_t fmm_reserve(FmmWorkspace& w, int64_t n, int, int, int, std::string& err) {
    return fmm_reserve_particles(w, n, err);
}

template <int P>
struct FmmPasses {
    using Ops = FmmOps<P>;

    static cudaError_t upward(FmmWorkspace& w, const std::vector<int>& lvl, cudaStream_t st, uint64_t& launches) {
        cudaError_t e;
        std::string err;
        if ((e = fmm_grow(w.M, w.ml_doubles_M, (size_t)w.ncells * 3 * Ops::NM, err)) != cudaSuccess) return e;
        if ((e = fmm_grow(w.L, w.ml_doubles_L, (size_t)w.ncells * 3 * Ops::NL, err)) != cudaSuccess) return e;
        if ((e = cudaMemsetAsync(w.M, 0, sizeof(double) * (size_t)w.ncells * 3 * Ops::NM, st)) != cudaSuccess) return e;
        fmm_p2m_kernel<P><<<(w.ncells + 3) / 4, 128, 0, st>>>(w.cells, w.ncells, w.rec, w.M);
        ++launches;
        for (int l = (int)lvl.size() - 2; l >= 0; --l) {   // lvl[l] .. lvl[l+1] is level l
            const int c0 = lvl[l], c1 = lvl[l + 1];
            if (l == (int)lvl.size() - 2) continue;          // deepest level has no children
            fmm_m2m_kernel<P><<<((c1 - c0) * 3 + 127) / 128, 128, 0, st>>>(w.cells, c0, c1, w.M);
            ++launches;
        }
        return cudaGetLastError();
    }

    static cudaError_t downward(FmmWorkspace& w, const std::vector<int>& lvl, cudaStream_t st, uint64_t& launches) {
        cudaError_t e;
        if ((e = cudaMemsetAsync(w.L, 0, sizeof(double) * (size_t)w.ncells * 3 * Ops::NL, st)) != cudaSuccess) return e;
        const size_t smem = sizeof(double) * 3 * Ops::NL * 32;
        if ((e = cudaFuncSetAttribute(fmm_m2l_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
            return e;
        const FmmCell* cells = w.cells_eval ? w.cells_eval : w.cells;
        const double* M = w.M_eval ? w.M_eval : w.M;
        fmm_m2l_kernel<P><<<w.ncells, 32, smem, st>>>(cells, w.ncells, w.m2l_sorted, w.m2l_off, M, w.L, w.halo);
        ++launches;
        for (int l = 1; l + 1 < (int)lvl.size(); ++l) {
            const int c0 = lvl[l], c1 = lvl[l + 1];
            fmm_l2l_kernel<P><<<((c1 - c0) * 3 + 127) / 128, 128, 0, st>>>(cells, c0, c1, w.L);
            ++launches;
        }
        return cudaGetLastError();
    }

    template <int KERNEL, int REP>
    static constexpr size_t leaves_uj_smem() {
        return sizeof(double) * ((KERNEL == K_GAUSSIANERF ? (size_t)VPM_GG_DOUBLES * REP : 0) +
                                 LeafGeom<REP>::WARPS * (((3 * Ops::NL + 1) & ~1) + (size_t)2 * LeafGeom<REP>::BATCH * REC_REALS));
    }

    template <int KERNEL, int REP>
    static cudaError_t leaves_uj_launch(FmmWorkspace& w, const double* gh_table, cudaStream_t st, uint64_t& launches) {
        constexpr int WARPS = LeafGeom<REP>::WARPS;
        const size_t smem = leaves_uj_smem<KERNEL, REP>();
        auto kfn = fmm_leaf_uj_kernel<KERNEL, P, REP>;
        cudaError_t e = cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int nl = w.leaf_hi - w.leaf_lo;
        if (nl <= 0) return cudaSuccess;
        int grid = 0;
        if ((e = fmm_leaf_grid(w, kfn, 32 * WARPS, smem, (nl + WARPS - 1) / WARPS, st, grid)) != cudaSuccess) return e;
        kfn<<<grid, 32 * WARPS, smem, st>>>(
            w.cells_eval ? w.cells_eval : w.cells, w.leaves + w.leaf_lo, nl, &w.counters->next, w.runs, w.p2p_off, w.rec, w.sx, w.sy, w.sz, w.L, gh_table, w.sU, w.sJ, w.lds, w.halo);
        ++launches;
        return cudaGetLastError();
    }

    // `copies` = 8 asks for the bank-conflict-free replicated G table (gaussianerf only, and only for p <= 4: one 512-thread
    // CTA caps the kernel at 128 registers, which the L2P of p >= 5 overflows); anything else runs the single-copy kernel.
    template <int KERNEL>
    static cudaError_t leaves_uj(FmmWorkspace& w, int copies, const double* gh_table, cudaStream_t st, uint64_t& launches) {
        if (KERNEL == K_GAUSSIANERF && copies == 8 && P <= 4 && leaves_uj_smem<K_GAUSSIANERF, 8>() <= 227 * 1024)
            return leaves_uj_launch<K_GAUSSIANERF, 8>(w, gh_table, st, launches);
        return leaves_uj_launch<KERNEL, 1>(w, gh_table, st, launches);
    }

    static cudaError_t leaves_uj(FmmWorkspace& w, int kernel, int block, const double* gh_table, cudaStream_t st,
                                 uint64_t& launches) {
        switch (kernel) {
        case K_GAUSSIANERF: return leaves_uj<K_GAUSSIANERF>(w, block, gh_table, st, launches);
        case K_WINCKELMANS: return leaves_uj<K_WINCKELMANS>(w, block, gh_table, st, launches);
        case K_GAUSSIAN: return leaves_uj<K_GAUSSIAN>(w, block, gh_table, st, launches);
        default: return leaves_uj<K_SINGULAR>(w, block, gh_table, st, launches);
        }
    }
};

// Bounding cube + Morton keys + radix sort of the particles: fills w.keys (sorted), w.perm (Morton slot -> particle) and
// the root cell geometry.  One host read-back (the 6 bounds).
struct FmmRoot {
    double cx, cy, cz, side;
};

inline cudaError_t fmm_sort(FmmWorkspace& w, const double* soa, int64_t ld, int64_t n, cudaStream_t st, uint64_t& launches,
                            std::string& err, FmmRoot* root_out = nullptr) {
    const double* X = soa + (size_t)F_X * ld;
    const double* Y = soa + (size_t)(F_X + 1) * ld;
    const double* Z = soa + (size_t)(F_X + 2) * ld;
    const int nb = 128;
    fmm_bounds_kernel<<<nb, 256, 0, st>>>(X, Y, Z, n, w.bounds);
    ++launches;
    std::vector<double> hb(6 * nb);
    FMM_TRY(cudaMemcpyAsync(hb.data(), w.bounds, sizeof(double) * 6 * nb, cudaMemcpyDeviceToHost, st));
    FMM_TRY(cudaStreamSynchronize(st));
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (int b = 0; b < nb; ++b)
        for (int c = 0; c < 3; ++c) {
            lo[c] = std::min(lo[c], hb[6 * b + 2 * c]);
            hi[c] = std::max(hi[c], -hb[6 * b + 2 * c + 1]);
        }
    for (int c = 0; c < 3; ++c)
        if (!(lo[c] <= hi[c]) || !std::isfinite(lo[c]) || !std::isfinite(hi[c])) {
            err = "particle positions are not finite";
            return cudaErrorInvalidValue;
        }
    double side = std::max(std::max(hi[0] - lo[0], hi[1] - lo[1]), hi[2] - lo[2]);
    side = side > 0 ? side * (1.0 + 1e-9) : 1.0;
    const double cx = 0.5 * (lo[0] + hi[0]), cy = 0.5 * (lo[1] + hi[1]), cz = 0.5 * (lo[2] + hi[2]);
    const double x0 = cx - 0.5 * side, y0 = cy - 0.5 * side, z0 = cz - 0.5 * side;
    const unsigned nbk = (unsigned)((n + 255) / 256);
    fmm_keys_kernel<<<nbk, 256, 0, st>>>(X, Y, Z, n, x0, y0, z0, 2097152.0 / side, w.keys, w.perm);
    ++launches;
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, w.keys, w.keys_alt, w.perm, w.perm_alt, (int)n, 0, 63, st);
    FMM_TRY(fmm_cub(w, tb, err));
    tb = w.cub_bytes;
    FMM_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys, w.keys_alt, w.perm, w.perm_alt, (int)n, 0, 63, st));
    ++launches;
    std::swap(w.keys, w.keys_alt);
    std::swap(w.perm, w.perm_alt);
    if (root_out) *root_out = FmmRoot{cx, cy, cz, side};
    return cudaSuccess;
}

#define FMM_CUB(call_with_tmp)                                              \
    do {                                                                     \
        size_t tb = 0;                                                       \
        void* tmp = nullptr;                                                 \
        (void)(call_with_tmp);                                               \
        FMM_TRY(fmm_cub(w, tb, err));                                        \
        tmp = w.cub_tmp;                                                     \
        tb = w.cub_bytes;                                                    \
        FMM_TRY(call_with_tmp);                                              \
    } while (0)

// The adaptive octree over the Morton-sorted keys in w.keys (n particles), level by level; fills w.cells, w.ncells, w.nlevels,
// lvl (cell index range of every level) and the leaf list (w.leaves, w.nleaves).  hpre / Lc: fmm_cell_splits.
inline cudaError_t fmm_tree(FmmWorkspace& w, int64_t n, int ncrit, const FmmRoot& rt, std::vector<int>& lvl, cudaStream_t st,
                            uint64_t& launches, std::string& err, const int* hpre = nullptr, int Lc = 0) {
    fmm_tic(w, 0, st);
    // the cell arrays start from an estimate and the build restarts if they are too small; the estimate follows the CURRENT
    // particle count (a field that grew from a small first call must not rely on doublings alone)
    FMM_TRY(fmm_reserve_cells(w, std::max<int64_t>(8192, 6 * n / std::max(ncrit, 1)), err));
    int ncells = 1;
    for (int attempt = 0;; ++attempt) {
        FmmCell root;
        root.start = 0; root.count = (int)n; root.parent = -1; root.child0 = -1; root.nchild = 0; root.level = 0;
        root.cx = rt.cx; root.cy = rt.cy; root.cz = rt.cz; root.R = 0.5 * rt.side; root.smax = 0.0; root.pad_ = 0.0;
        FMM_TRY(cudaMemcpyAsync(w.cells, &root, sizeof(root), cudaMemcpyHostToDevice, st));
        lvl.clear();
        lvl.push_back(0);
        lvl.push_back(1);
        ncells = 1;
        bool overflow = false;
        while (true) {
            const int c0 = lvl[lvl.size() - 2], c1 = lvl.back();
            const int nc = c1 - c0;
            fmm_split_count_kernel<<<(nc + 127) / 128, 128, 0, st>>>(w.cells, c0, c1, w.keys, ncrit, w.nchild, hpre, Lc, w.rec);
            FMM_CUB(cub::DeviceScan::ExclusiveSum(tmp, tb, w.nchild, w.child_off, nc, st));
            int last_off = 0, last_n = 0;
            FMM_TRY(cudaMemcpyAsync(&last_off, w.child_off + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            FMM_TRY(cudaMemcpyAsync(&last_n, w.nchild + nc - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
            FMM_TRY(cudaStreamSynchronize(st));
            const int nnew = last_off + last_n;
            launches += 2;
            if ((int64_t)ncells + nnew > w.cap_cells) {
                overflow = true;
                break;
            }
            fmm_split_emit_kernel<<<(nc + 127) / 128, 128, 0, st>>>(w.cells, c0, c1, w.keys, ncrit, w.child_off, ncells, hpre, Lc, w.rec);
            ++launches;
            if (nnew == 0) break;
            ncells += nnew;
            lvl.push_back(ncells);
        }
        if (!overflow) break;
        // single-child chains are not compressed, so the only hard bound is one cell per particle per level
        if (attempt >= 12 || (int64_t)w.cap_cells > n * (FMM_MAXLEVEL + 1) + 64) {
            err = "FMM: octree needs more cells than particles allow (corrupt positions?)";
            return cudaErrorMemoryAllocation;
        }
        FMM_TRY(fmm_reserve_cells(w, (int64_t)w.cap_cells * 2, err));
    }
    w.ncells = ncells;
    w.nlevels = (int)lvl.size() - 1;
    // ---- leaves
    fmm_mark_leaves_kernel<<<(ncells + 255) / 256, 256, 0, st>>>(w.cells, ncells, w.leaf_flag);
    FMM_CUB(cub::DeviceScan::ExclusiveSum(tmp, tb, w.leaf_flag, w.leaf_pos, ncells, st));
    fmm_collect_leaves_kernel<<<(ncells + 255) / 256, 256, 0, st>>>(w.leaf_flag, w.leaf_pos, ncells, w.leaves);
    launches += 3;
    int lp = 0, lf = 0;
    FMM_TRY(cudaMemcpyAsync(&lp, w.leaf_pos + ncells - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FMM_TRY(cudaMemcpyAsync(&lf, w.leaf_flag + ncells - 1, sizeof(int), cudaMemcpyDeviceToHost, st));
    FMM_TRY(cudaStreamSynchronize(st));
    w.nleaves = lp + lf;
    w.leaf_lo = 0;
    w.leaf_hi = w.nleaves;
    fmm_toc(w, 0, st);
    return cudaSuccess;
}

// smax (largest core size per cell) of the cells in `cells`, leaves first then level by level upward.
inline void fmm_smax(FmmWorkspace& w, FmmCell* cells, int ncells, const std::vector<int>& lvl, cudaStream_t st, uint64_t& launches) {
    fmm_smax_leaf_kernel<<<(ncells + 127) / 128, 128, 0, st>>>(cells, ncells, w.rec);
    ++launches;
    for (int l = (int)lvl.size() - 3; l >= 0; --l) {
        fmm_smax_up_kernel<<<(lvl[l + 1] - lvl[l] + 127) / 128, 128, 0, st>>>(cells, lvl[l], lvl[l + 1]);
        ++launches;
    }
}

// Dual tree traversal from the seed (target, source) pairs over the cell array `cells` (targets are cells [0, ntarget); sources
// may sit anywhere in `cells` — the local essential tree appends the other ranks' trees behind the rank's own), then the
// lists sorted by (target, source), their per-target offsets and the particle run of every P2P entry.  Pair buffers start
// from an estimate, grow and start over if a list overflows.  count_at must already hold the particle count of every
// SOURCE leaf outside [0, ntarget) (remote leaves), indexed by first particle; own leaves are filled in here.
inline cudaError_t fmm_lists(FmmWorkspace& w, const FmmCell* cells, int ntarget, double theta, double nzs_factor, const int* mine,
                             const uint64_t* seeds, int nseeds, int nparts, cudaStream_t st, uint64_t& launches, std::string& err) {
    fmm_tic(w, 1, st);
    if (!w.front_a) {
        w.cap_pairs = (unsigned int)std::min<int64_t>(std::max<int64_t>(1 << 20, 192LL * ntarget / nparts), 1500000000LL);
        w.cap_p2p = (unsigned int)std::min<int64_t>(std::max<int64_t>(1 << 20, 96LL * ntarget / nparts), 1500000000LL);
        FMM_TRY(fmm_alloc_pairs(w, err));
    }
    FmmCounters zero = {0, 0, 0, 0};
    FmmCounters hc = zero;
    for (int attempt = 0;; ++attempt) {
        FMM_TRY(cudaMemcpyAsync(w.counters, &zero, sizeof(zero), cudaMemcpyHostToDevice, st));
        FMM_TRY(cudaMemcpyAsync(w.front_a, seeds, sizeof(uint64_t) * nseeds, cudaMemcpyHostToDevice, st));
        unsigned int nfront = (unsigned int)nseeds;
        uint64_t *fa = w.front_a, *fb = w.front_b;
        hc = zero;
        while (nfront > 0) {
            fmm_traverse_kernel<<<(nfront + 255) / 256, 256, 0, st>>>(cells, fa, nfront, theta, nzs_factor, fb, w.cap_pairs, w.m2l,
                                                                     w.cap_pairs, w.p2p, w.cap_p2p, w.counters, mine);
            ++launches;
            FMM_TRY(cudaMemcpyAsync(&hc, w.counters, sizeof(hc), cudaMemcpyDeviceToHost, st));
            FMM_TRY(cudaStreamSynchronize(st));
            if (hc.overflow) break;
            nfront = hc.next;
            const unsigned int z = 0;
            FMM_TRY(cudaMemcpyAsync(&w.counters->next, &z, sizeof(z), cudaMemcpyHostToDevice, st));
            std::swap(fa, fb);
        }
        if (!hc.overflow) break;
        if (attempt >= 8 || w.cap_pairs > 1400000000u || w.cap_p2p > 1400000000u) {
            err = "FMM: interaction lists overflowed the workspace";
            return cudaErrorMemoryAllocation;
        }
        // counters keep counting past the capacity, so they say which list to grow
        if (hc.p2p >= w.cap_p2p) w.cap_p2p = (unsigned int)std::min<uint64_t>(std::max<uint64_t>((uint64_t)w.cap_p2p * 2, (uint64_t)hc.p2p * 3 / 2), 1500000000ULL);
        if (hc.m2l >= w.cap_pairs || hc.next >= w.cap_pairs)
            w.cap_pairs = (unsigned int)std::min<uint64_t>((uint64_t)w.cap_pairs * 2, 1500000000ULL);
        FMM_TRY(fmm_alloc_pairs(w, err));
    }
    w.n_m2l = hc.m2l;
    w.n_p2p = hc.p2p;
    // ---- sort the lists by (target, source) and index them per target cell
    int bits = 1;
    while ((1 << bits) < ntarget) ++bits;
    if (w.n_m2l > 0) {
        FMM_CUB(cub::DeviceRadixSort::SortKeys(tmp, tb, w.m2l, w.m2l_sorted, (int)w.n_m2l, 0, 32 + bits, st));
        ++launches;
    }
    if (w.n_p2p > 0) {
        FMM_CUB(cub::DeviceRadixSort::SortKeys(tmp, tb, w.p2p, w.p2p_sorted, (int)w.n_p2p, 0, 32 + bits, st));
        ++launches;
    }
    fmm_list_offsets_kernel<<<(ntarget + 256) / 256, 256, 0, st>>>(w.m2l_sorted, w.n_m2l, ntarget, w.m2l_off);
    fmm_list_offsets_kernel<<<(ntarget + 256) / 256, 256, 0, st>>>(w.p2p_sorted, w.n_p2p, ntarget, w.p2p_off);
    fmm_leaf_counts_kernel<<<(w.nleaves + 255) / 256, 256, 0, st>>>(cells, w.leaves, w.nleaves, w.count_at);
    if (w.n_p2p > 0) fmm_p2p_runs_kernel<<<(w.n_p2p + 255) / 256, 256, 0, st>>>(w.p2p_sorted, w.n_p2p, w.count_at, w.runs);
    launches += 4;
    FMM_TRY(cudaGetLastError());
    fmm_toc(w, 1, st);
    return cudaSuccess;
}

// Build the adaptive octree and the interaction lists for the particles of the field.
inline cudaError_t fmm_build(FmmWorkspace& w, const double* soa, int64_t ld, int64_t n, int ncrit, double theta,
                             double nzs_factor, std::vector<int>& lvl, cudaStream_t st, uint64_t& launches, std::string& err,
                             int part = 0, int nparts = 1) {
    FMM_TRY(fmm_reserve_particles(w, n, err));
    FmmRoot rt;
    {
        cudaError_t e0 = fmm_sort(w, soa, ld, n, st, launches, err, &rt);
        if (e0 != cudaSuccess) return e0;
    }
    const unsigned nbk = (unsigned)((n + 255) / 256);
    fmm_gather_kernel<<<nbk, 256, 0, st>>>(soa, ld, n, w.perm, w.sx, w.sy, w.sz, w.rec);
    ++launches;
    {
        cudaError_t e1 = fmm_tree(w, n, ncrit, rt, lvl, st, launches, err);
        if (e1 != cudaSuccess) return e1;
    }
    const int ncells = w.ncells;
    const int* mine = nullptr;
    if (nparts > 1) {
        // leaves in Morton order; this rank takes the leaves covering particles [part, part + 1) * n / nparts of that order
        // and marks them and their ancestors: only those target cells are traversed / receive M2L, L2L, L2P, P2P
        fmm_leaf_starts_kernel<<<(w.nleaves + 255) / 256, 256, 0, st>>>(w.cells, w.leaves, w.nleaves, w.leaf_keys);
        FMM_CUB(cub::DeviceRadixSort::SortPairs(tmp, tb, w.leaf_keys, w.leaf_keys_alt, w.leaves, w.leaves_alt, w.nleaves, 0, 32, st));
        std::swap(w.leaf_keys, w.leaf_keys_alt);
        std::swap(w.leaves, w.leaves_alt);
        const int p_lo = (int)(n * part / nparts), p_hi = (int)(n * (part + 1) / nparts);
        fmm_leaf_range_kernel<<<1, 32, 0, st>>>(w.leaf_keys, w.nleaves, p_lo, p_hi, w.leaf_flag);
        int range[2] = {0, 0};
        FMM_TRY(cudaMemcpyAsync(range, w.leaf_flag, sizeof(range), cudaMemcpyDeviceToHost, st));
        FMM_TRY(cudaStreamSynchronize(st));
        w.leaf_lo = range[0];
        w.leaf_hi = part == nparts - 1 ? w.nleaves : range[1];
        FMM_TRY(cudaMemsetAsync(w.mine, 0, sizeof(int) * ncells, st));
        if (w.leaf_hi > w.leaf_lo)
            fmm_flag_owned_kernel<<<(w.leaf_hi - w.leaf_lo + 255) / 256, 256, 0, st>>>(w.cells, w.leaves + w.leaf_lo, w.leaf_hi - w.leaf_lo, w.mine);
        launches += 4;
        mine = w.mine;
    }
    if (nzs_factor > 0.0) fmm_smax(w, w.cells, ncells, lvl, st, launches);
    const uint64_t rootpair = 0;
    return fmm_lists(w, w.cells, ncells, theta, nzs_factor, mine, &rootpair, 1, nparts, st, launches, err);
}

// Re-gather the Morton-ordered source records from the state (same positions and strengths, new core sizes): what an
// evaluation needs when the tree, the lists and the far field of the previous one are still valid (fmm_evaluate far_valid).
inline cudaError_t fmm_regather(FmmWorkspace& w, const double* soa, int64_t ld, int64_t n, cudaStream_t st, uint64_t& launches) {
    fmm_gather_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(soa, ld, n, w.perm, w.sx, w.sy, w.sz, w.rec);
    ++launches;
    return cudaGetLastError();
}

// far_valid: the local expansions L of the previous evaluation are still right (same positions, strengths, tree and lists;
// the far field is the singular kernel, so it does not depend on sigma): skip P2M/M2M/M2L/L2L and only redo L2P + near field.
template <int P>
inline cudaError_t fmm_evaluate_p(FmmWorkspace& w, int kernel, int block, const double* gh_table, const std::vector<int>& lvl,
                                  cudaStream_t st, uint64_t& launches, bool far_valid, bool skip_upward, int stage) {
    cudaError_t e;
    if (!far_valid && stage != 2) {
        if (!skip_upward) {
            fmm_tic(w, 2, st);
            if ((e = FmmPasses<P>::upward(w, lvl, st, launches)) != cudaSuccess) return e;
            fmm_toc(w, 2, st);
        }
        fmm_tic(w, 3, st);
        if ((e = FmmPasses<P>::downward(w, lvl, st, launches)) != cudaSuccess) return e;
        fmm_toc(w, 3, st);
    }
    if (stage == 1) return cudaSuccess;
    fmm_tic(w, 4, st);
    e = FmmPasses<P>::leaves_uj(w, kernel, block, gh_table, st, launches);
    fmm_toc(w, 4, st);
    return e;
}

inline cudaError_t fmm_evaluate(FmmWorkspace& w, int p, int kernel, int block, const double* gh_table,
                                const std::vector<int>& lvl, cudaStream_t st, uint64_t& launches, bool far_valid = false,
                                bool skip_upward = false, int stage = 0) {
    switch (p) {
    case 2: return fmm_evaluate_p<2>(w, kernel, block, gh_table, lvl, st, launches, far_valid, skip_upward, stage);
    case 3: return fmm_evaluate_p<3>(w, kernel, block, gh_table, lvl, st, launches, far_valid, skip_upward, stage);
    case 4: return fmm_evaluate_p<4>(w, kernel, block, gh_table, lvl, st, launches, far_valid, skip_upward, stage);
    case 5: return fmm_evaluate_p<5>(w, kernel, block, gh_table, lvl, st, launches, far_valid, skip_upward, stage);
    case 6: return fmm_evaluate_p<6>(w, kernel, block, gh_table, lvl, st, launches, far_valid, skip_upward, stage);
    default: return cudaErrorInvalidValue;
    }
}

inline cudaError_t fmm_estr(FmmWorkspace& w, int kernel, int block, int transposed, const double* z_table, cudaStream_t st,
                            uint64_t& launches) {
    (void)block;
    const int nl = w.leaf_hi - w.leaf_lo;
    if (nl <= 0) return cudaSuccess;
    const size_t smem = sizeof(double) * ((kernel == K_GAUSSIANERF ? ((VPM_GZ_NINT + 1) & ~1) : 0) +
                                          LEAF_WARPS * (size_t)2 * LEAF_BATCH * REC_REALS);
    int grid = 0;
    cudaError_t eg = cudaSuccess;
    fmm_tic(w, 5, st);
#define FMM_ESTR_CASE(K)                                                                                                   \
    cudaFuncSetAttribute(fmm_leaf_estr_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
    if ((eg = fmm_leaf_grid(w, fmm_leaf_estr_kernel<K>, 32 * LEAF_WARPS, smem, (nl + LEAF_WARPS - 1) / LEAF_WARPS, st, grid)) != cudaSuccess) return eg; \
    fmm_leaf_estr_kernel<K><<<grid, 32 * LEAF_WARPS, smem, st>>>(                                                          \
        w.cells_eval ? w.cells_eval : w.cells, w.leaves + w.leaf_lo, nl, &w.counters->next, w.runs, w.p2p_off, w.rec, w.sx, w.sy, w.sz, w.sJ, w.lds, transposed, z_table, w.sE, w.halo)
    switch (kernel) {
    case K_GAUSSIANERF: FMM_ESTR_CASE(K_GAUSSIANERF); break;
    case K_WINCKELMANS: FMM_ESTR_CASE(K_WINCKELMANS); break;
    case K_GAUSSIAN: FMM_ESTR_CASE(K_GAUSSIAN); break;
    default: FMM_ESTR_CASE(K_SINGULAR); break;
    }
#undef FMM_ESTR_CASE
    ++launches;
    fmm_toc(w, 5, st);
    return cudaGetLastError();
}

}  // namespace vpm
