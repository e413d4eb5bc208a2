// fmm_let.cuh — multi-GPU UJ_fmm with a LOCAL ESSENTIAL TREE: the per-rank phases.
//
// north_star: "the FMM path shards subtrees with a local-essential-tree exchange".  The reference has no distributed mode
// (it is a single Julia process, /root/reference/src/FLOWUnsteady_simulation.jl:339-447), so ownership is ours to define as
// long as U, J (and E_str) come back in the caller's particle order.  One evaluation, seen from rank r of G:
//
//   1 bounds     local min/max of the positions                      -> all-reduce (6 doubles): the GLOBAL root cube
//   2 keys       63-bit Morton keys of the HOME particles (the ones whose state lives here), radix sort, level-Lc
//                histogram (8^Lc bins, Morton order)                 -> all-reduce: the global histogram
//   3 partition  every rank cuts the Morton curve into G ranges of (nearly) equal particle count at the SAME unit
//                boundaries (a unit = a level-Lc bin, or a coarser cell that is a leaf of the global tree), so every subtree
//                below the top has exactly one owner               -> all-to-all of 7-double particle rows to the owners
//   4 build      the owner sorts what it received and builds its part of the GLOBAL octree: cells above level Lc are split
//                by the GLOBAL count (fmm_cell_splits), cells below by the local one — the union of the ranks' trees IS the
//                one-GPU tree, a top cell simply appears on every rank that owns particles in it, with a partial multipole
//   5 upward     P2M / M2M on the owner's cells
//   6 exchange   all-gather of the tree skeletons (72 B per cell), of the multipoles and of the source records: the essential
//                data of the other ranks is appended behind the rank's own in ONE cell / multipole / record array
//   7 evaluate   dual traversal of the rank's OWN target cells against every rank's tree (G seed pairs), M2L, L2L, L2P and
//                the near field with the single-GPU kernels: sources are just indices into the combined arrays
//   8 return     rows of U, J (E_str after a second near-field pass over exchanged E_str records) go back to the home ranks
//                with the inverse all-to-all and are scattered into the state
//
// Linear in the sources, the sum over the ranks' partial top cells equals the one-GPU M2L of the whole cell, and every
// other pair is the same pair: results match one GPU to summation round-off (tests/test_let.py, 1e-12).
// The collectives themselves live in flowunsteady_b200/dist.py (NCCL via torch.distributed — plumbing); everything here is
// per-rank device work exposed through include/vpmb200.h (vpmb200_let_*).
#pragma once

#include "fmm_host.cuh"

namespace vpm {

constexpr int LET_ROW = 7;            // exchanged particle row: x, y, z, Gamma(3), sigma
constexpr int LET_MAX_LC = 6;         // 8^6 = 262,144 bins at most

struct FmmLet {
    // home side (particles whose state lives on this rank)
    uint64_t *hkeys = nullptr, *hkeys_alt = nullptr;
    int *hperm = nullptr, *hperm_alt = nullptr;
    int64_t hcap = 0, n_home = 0;
    int* hist = nullptr;               // level-Lc histogram (all-reduced in place by the caller), then its prefix sum
    int* hpre = nullptr;               // exclusive prefix sum of the global histogram, bins + 1 entries
    double* binmax = nullptr;          // per-bin largest sigma (nonzero_sigma), all-reduced (max) by the caller
    int* split_idx = nullptr;          // device scratch: positions of the splitters in the sorted home keys
    int Lc = 0, bins = 0;
    FmmRoot cube{0, 0, 0, 1};
    int nparts = 1, part = 0;
    std::vector<uint64_t> splitters;   // nparts + 1 key bounds
    std::vector<int64_t> send_counts;
    // work-weighted cut: interaction work counted per level-Lc bin in the previous evaluation (let_count_work), all-reduced
    // in place by the caller before the next let_partition
    long long* work = nullptr;
    // owner side
    const double* rows = nullptr;      // received rows (n_own x 7), owned by the caller, alive until the evaluation ends
    int64_t n_own = 0, n_all = 0;
    int ncells_own = 0, ncells_all = 0;
    FmmCell* cells_all = nullptr;      // [own cells | rank 0's | rank 1's | ...] (own rank skipped), indices fixed up
    int64_t cap_cells_all = 0;
    double* M_all = nullptr;
    size_t cap_M_all = 0;
    std::vector<int> lvl;              // own tree levels
    std::vector<int> cell_off;         // offset of every rank's tree inside cells_all (own rank: 0)
    std::vector<int64_t> part_off;     // offset of every rank's particles inside w.rec (own rank: 0)
    std::vector<int64_t> ncells_of;    // cells of every rank's tree
    bool far_valid = false;            // L of the last evaluation still valid (DynamicSFS second evaluation)
    int P = 0;
    // demand-driven halo ("let_halo"): only the skeletons are gathered; multipoles and records of the cells / leaves the
    // traversal actually touches are requested from their owners
    bool halo_mode = false;
    int nleaves_remote = 0;            // leaves of the other ranks' trees (their ordinals index rleaf / pneed / poff)
    std::vector<int> leaf_off;         // first ordinal of every rank's leaves (own rank: 0, unused)
    std::vector<int64_t> nleaves_of;
    int2* rleaf = nullptr;             // remote leaf ordinal -> (first record in the OWNER's order, count)
    int *mneed = nullptr, *mslot = nullptr, *pneed = nullptr, *pslot = nullptr, *pcnt = nullptr, *poff = nullptr;
    int* req_cells = nullptr;          // needed remote cells as owner-local ids, grouped by owner
    int2* req_leaf = nullptr;          // needed remote leaves as (first record, count) in the owner's order, grouped by owner
    int64_t cap_hc = 0, cap_hl = 0;    // capacities of the per-cell / per-remote-leaf arrays
    int* serve_off = nullptr;          // owner side: record offset of every requested leaf
    int64_t cap_serve = 0;
};

inline void let_free(FmmLet& t) {
    void* ptrs[] = {t.hkeys, t.hkeys_alt, t.hperm, t.hperm_alt, t.hist, t.hpre, t.binmax, t.split_idx, t.cells_all, t.M_all, t.work,
                    t.rleaf, t.mneed, t.mslot, t.pneed, t.pslot, t.pcnt, t.poff, t.req_cells, t.req_leaf, t.serve_off};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    t = FmmLet();
}

// ---- kernels -----------------------------------------------------------------------------------------------------------
// hist[b] = number of sorted keys in Morton bin b of level Lc (two binary searches per bin)
__global__ void let_hist_kernel(const uint64_t* __restrict__ keys, int n, int Lc, int bins, int* __restrict__ hist) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bins) return;
    const int shift = 3 * (FMM_MAXLEVEL - Lc);
    const int lo = lower_bound_key(keys, 0, n, (uint64_t)b << shift);
    const int hi = b + 1 == bins ? n : lower_bound_key(keys, lo, n, (uint64_t)(b + 1) << shift);
    hist[b] = hi - lo;
}

// per-bin largest sigma of the HOME particles (for the sigma-aware acceptance of nonzero_sigma = true)
__global__ void let_binmax_kernel(const uint64_t* __restrict__ keys, const int* __restrict__ perm, const double* __restrict__ sigma,
                                  int n, int Lc, int bins, double* __restrict__ binmax) {
    int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= bins) return;
    const int shift = 3 * (FMM_MAXLEVEL - Lc);
    const int lo = lower_bound_key(keys, 0, n, (uint64_t)b << shift);
    const int hi = b + 1 == bins ? n : lower_bound_key(keys, lo, n, (uint64_t)(b + 1) << shift);
    double m = 0.0;
    for (int i = lo; i < hi; ++i) m = fmax(m, sigma[perm[i]]);
    binmax[b] = m;
}

__global__ void let_find_splits_kernel(const uint64_t* __restrict__ keys, int n, const uint64_t* __restrict__ split, int ns,
                                       int* __restrict__ idx) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < ns) idx[k] = lower_bound_key(keys, 0, n, split[k]);
}

// rows[i] = (x, y, z, Gamma, sigma) of home particle perm[i]  (Morton order: each destination's rows are contiguous)
__global__ void let_pack_rows_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, const int* __restrict__ perm,
                                     double* __restrict__ rows) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = perm[i];
    double* r = rows + i * LET_ROW;
#pragma unroll
    for (int c = 0; c < LET_ROW; ++c) r[c] = soa[(size_t)c * ld + p];   // rows 0..6 of the state are X, Gamma, sigma
}

__global__ void let_keys_rows_kernel(const double* __restrict__ rows, int64_t n, double x0, double y0, double z0, double inv_cell,
                                     uint64_t* __restrict__ keys, int* __restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lim = 2097151.0;
    const double* r = rows + i * LET_ROW;
    uint64_t ix = (uint64_t)fmin(fmax((r[0] - x0) * inv_cell, 0.0), lim);
    uint64_t iy = (uint64_t)fmin(fmax((r[1] - y0) * inv_cell, 0.0), lim);
    uint64_t iz = (uint64_t)fmin(fmax((r[2] - z0) * inv_cell, 0.0), lim);
    keys[i] = (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
    perm[i] = (int)i;
}

// Morton-ordered targets + UJ source records of the owner's particles, from the received rows (cf. fmm_gather_kernel)
__global__ void let_gather_rows_kernel(const double* __restrict__ rows, int64_t n, const int* __restrict__ perm,
                                       double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz,
                                       double* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* q = rows + (int64_t)perm[i] * LET_ROW;
    const double x = q[0], y = q[1], z = q[2], gx = q[3], gy = q[4], gz = q[5], sg = q[6];
    double si = 1.0 / sg, si2 = si * si, si3 = si2 * si;
    sx[i] = x; sy[i] = y; sz[i] = z;
    double2* r = reinterpret_cast<double2*>(rec + (size_t)i * REC_REALS);
    r[0] = make_double2(x, y);
    r[1] = make_double2(z, -CONST4 * gx);
    r[2] = make_double2(-CONST4 * gy, -CONST4 * gz);
    r[3] = make_double2(VPM_GT_TFAR * (sg * sg), si3);
    r[4] = make_double2(si3 * si2, si2);
}

// E_str records of the owner's particles (estr_direct.cuh layout) from the received rows and the Morton-ordered TOTAL J
__global__ void let_estr_records_kernel(const double* __restrict__ rows, int64_t n, const int* __restrict__ perm,
                                        const double* __restrict__ sJ, int64_t ldj, int transposed, double zeta_norm,
                                        double* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* q = rows + (int64_t)perm[i] * LET_ROW;
    const double x = q[0], y = q[1], z = q[2], g0 = q[3], g1 = q[4], g2 = q[5], sg = q[6];
    double Jq[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) Jq[c] = sJ[(size_t)c * ldj + i];
    double v0, v1, v2;
    if (transposed) {
        v0 = Jq[0] * g0 + Jq[1] * g1 + Jq[2] * g2;
        v1 = Jq[3] * g0 + Jq[4] * g1 + Jq[5] * g2;
        v2 = Jq[6] * g0 + Jq[7] * g1 + Jq[8] * g2;
    } else {
        v0 = Jq[0] * g0 + Jq[3] * g1 + Jq[6] * g2;
        v1 = Jq[1] * g0 + Jq[4] * g1 + Jq[7] * g2;
        v2 = Jq[2] * g0 + Jq[5] * g1 + Jq[8] * g2;
    }
    double si = 1.0 / sg, si2 = si * si;
    double c = zeta_norm * (si2 * si);
    double2* r = reinterpret_cast<double2*>(rec + (size_t)i * REC_REALS);
    r[0] = make_double2(x, y);
    r[1] = make_double2(z, si2);
    r[2] = make_double2(c * g0, c * g1);
    r[3] = make_double2(c * g2, c * v0);
    r[4] = make_double2(c * v1, c * v2);
}

// top cells (level < Lc): smax = largest sigma over ALL ranks' particles in the cell (range maximum over the global bins)
__global__ void let_top_smax_kernel(FmmCell* __restrict__ cells, int ncells, const uint64_t* __restrict__ keys, int Lc,
                                    const double* __restrict__ binmax) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const FmmCell cell = cells[c];
    if (cell.level >= Lc) return;
    const uint64_t q = keys[cell.start] >> (3 * (FMM_MAXLEVEL - cell.level));
    const int sh = 3 * (Lc - cell.level);
    double m = 0.0;
    for (uint64_t b = q << sh; b < ((q + 1) << sh); ++b) m = fmax(m, binmax[b]);
    cells[c].smax = m;
}

// another rank's tree appended behind the own one: cell and particle indices move by the block offsets
__global__ void let_attach_cells_kernel(const FmmCell* __restrict__ src, int nc, FmmCell* __restrict__ dst, int cell_off,
                                        int part_off) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    FmmCell v = src[c];
    v.start += part_off;
    if (v.parent >= 0) v.parent += cell_off;
    if (v.child0 >= 0) v.child0 += cell_off;
    dst[c] = v;
}

// count_at[first particle] = count for the leaves among cells [c0, c1)
__global__ void let_leaf_counts_kernel(const FmmCell* __restrict__ cells, int c0, int c1, int* __restrict__ count_at) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const FmmCell v = cells[c];
    if (v.nchild == 0) count_at[v.start] = v.count;
}

// Interaction work of the owner's cells, counted per level-Lc Morton bin (integer atomics: the sums are exact, so every rank
// cuts the next partition from identical numbers).  A leaf's near field costs (its particles) x (particles of its P2P list),
// twice with the E_str pass over the same pairs; an M2L costs about as much as LET_M2L_WORK particle pairs (measured: 0.25 ns
// per M2L against 0.015 ns per pair at p = 4).  Cells above level Lc (a handful) are not counted.
constexpr long long LET_M2L_WORK = 16;
__global__ void let_count_work_kernel(const FmmCell* __restrict__ cells, int ncells, const unsigned int* __restrict__ p2p_off,
                                      const int2* __restrict__ runs, const unsigned int* __restrict__ m2l_off,
                                      const uint64_t* __restrict__ keys, int Lc, long long* __restrict__ work) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const FmmCell cell = cells[c];
    if (cell.level < Lc) return;
    long long w = LET_M2L_WORK * (long long)(m2l_off[c + 1] - m2l_off[c]);
    if (cell.nchild == 0) {
        long long srcs = 0;
        for (unsigned int k = p2p_off[c]; k < p2p_off[c + 1]; ++k) srcs += runs[k].y;
        w += (long long)cell.count * srcs;
    }
    if (w > 0) atomicAdd(reinterpret_cast<unsigned long long*>(work) + (keys[cell.start] >> (3 * (FMM_MAXLEVEL - Lc))), (unsigned long long)w);
}

// Morton order -> the order the rows arrived in: out[perm[i]][k] = src[k][i]
__global__ void let_out_rows_kernel(const double* __restrict__ sA, int na, const double* __restrict__ sB, int nb, int64_t lds,
                                    int64_t n, const int* __restrict__ perm, double* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double* o = out + (int64_t)perm[i] * (na + nb);
    for (int k = 0; k < na; ++k) o[k] = sA[(size_t)k * lds + i];
    for (int k = 0; k < nb; ++k) o[na + k] = sB[(size_t)k * lds + i];
}

// results back on the home rank: res[i] belongs to home particle perm[i]; rows dst0.. of the state (+)= res columns
__global__ void let_finish_kernel(const double* __restrict__ res, int ncol, int64_t n, const int* __restrict__ perm,
                                  double* __restrict__ soa, int64_t ld, int rowA, int nA, int rowB, int nB, int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = perm[i];
    const double* r = res + i * ncol;
    for (int k = 0; k < nA; ++k) {
        double* d = soa + (size_t)(rowA + k) * ld + p;
        *d = accumulate ? *d + r[k] : r[k];
    }
    for (int k = 0; k < nB; ++k) {
        double* d = soa + (size_t)(rowB + k) * ld + p;
        *d = accumulate ? *d + r[nA + k] : r[nA + k];
    }
}

// ---- demand-driven halo ------------------------------------------------------------------------------------------------
// owner side, before the skeleton leaves: every leaf carries its ordinal among the tree's leaves (in the unused pad_ field)
__global__ void let_stamp_leaves_kernel(FmmCell* __restrict__ cells, int ncells, const int* __restrict__ leaf_pos) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) cells[c].pad_ = cells[c].nchild == 0 ? (double)leaf_pos[c] : -1.0;
}

// another rank's SKELETON behind the own cells: indices move as in let_attach_cells_kernel, but a leaf's particle range is not
// here — its `start` becomes n_own + (global ordinal of the remote leaf), the id the traversal pushes into the P2P list;
// rleaf keeps where the records sit on the owner, count_at (indexed by that id) the count
__global__ void let_attach_skeleton_kernel(const FmmCell* __restrict__ src, int nc, FmmCell* __restrict__ dst, int cell_off,
                                           int leaf_off, int n_own, int2* __restrict__ rleaf, int* __restrict__ count_at) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= nc) return;
    FmmCell v = src[c];
    if (v.parent >= 0) v.parent += cell_off;
    if (v.child0 >= 0) v.child0 += cell_off;
    if (v.nchild == 0) {
        const int o = leaf_off + (int)v.pad_;
        rleaf[o] = make_int2(v.start, v.count);
        count_at[n_own + o] = v.count;
        v.start = n_own + o;
    } else {
        v.start = n_own;
    }
    dst[c] = v;
}

__global__ void let_mark_m2l_kernel(const uint64_t* __restrict__ keys, unsigned int n, int cell_split, int* __restrict__ mneed) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int j = (int)(keys[k] & 0xffffffffu);
    if (j >= cell_split) mneed[j] = 1;
}
__global__ void let_mark_p2p_kernel(const uint64_t* __restrict__ keys, unsigned int n, int n_own, const int2* __restrict__ rleaf,
                                    int* __restrict__ pneed, int* __restrict__ pcnt) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int s = (int)(keys[k] & 0xffffffffu);
    if (s >= n_own) {
        pneed[s - n_own] = 1;
        pcnt[s - n_own] = rleaf[s - n_own].y;
    }
}
// needed cells of one owner's block -> owner-local ids at their compact slots
__global__ void let_compact_cells_kernel(const int* __restrict__ mneed, const int* __restrict__ mslot, int c0, int c1,
                                         int* __restrict__ req_cells) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c < c1 && mneed[c]) req_cells[mslot[c]] = c - c0;
}
__global__ void let_compact_leaves_kernel(const int* __restrict__ pneed, const int* __restrict__ pslot, const int2* __restrict__ rleaf,
                                          int n, int2* __restrict__ req_leaf) {
    int o = blockIdx.x * blockDim.x + threadIdx.x;
    if (o < n && pneed[o]) req_leaf[pslot[o]] = rleaf[o];
}
// P2P runs of remote leaves: (n_own + offset of the leaf's records inside the received halo buffer, count)
__global__ void let_halo_runs_kernel(const uint64_t* __restrict__ keys, unsigned int n, int n_own, const int2* __restrict__ rleaf,
                                     const int* __restrict__ poff, int2* __restrict__ runs) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int s = (int)(keys[k] & 0xffffffffu);
    if (s >= n_own) runs[k] = make_int2(n_own + poff[s - n_own], rleaf[s - n_own].y);
}
// owner side: requested multipoles, and the records of the requested leaves back to back (one warp per leaf)
__global__ void let_serve_M_kernel(const int* __restrict__ ids, int n, int nm3, const double* __restrict__ M, double* __restrict__ out) {
    int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)n * nm3) return;
    const int i = (int)(t / nm3), a = (int)(t % nm3);
    out[t] = M[(size_t)ids[i] * nm3 + a];
}
__global__ void let_serve_counts_kernel(const int2* __restrict__ leaves, int n, int* __restrict__ cnt) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) cnt[i] = leaves[i].y;
}
__global__ void let_serve_rec_kernel(const int2* __restrict__ leaves, const int* __restrict__ off, int n, const double* __restrict__ rec,
                                     double* __restrict__ out) {
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= n) return;
    const int2 l = leaves[i];
    const double* src = rec + (size_t)l.x * REC_REALS;
    double* dst = out + (size_t)off[i] * REC_REALS;
    for (int q = threadIdx.x & 31; q < l.y * REC_REALS; q += 32) dst[q] = src[q];
}

// ---- host side ---------------------------------------------------------------------------------------------------------
inline cudaError_t let_reserve_home(FmmLet& t, int64_t n, int Lc, std::string& err) {
    if (n > t.hcap || !t.hkeys) {
        void* ptrs[] = {t.hkeys, t.hkeys_alt, t.hperm, t.hperm_alt};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        t.hkeys = t.hkeys_alt = nullptr;
        t.hperm = t.hperm_alt = nullptr;
        const int64_t cap = std::max<int64_t>(n + n / 8, 4096);
        FMM_TRY(cudaMalloc(&t.hkeys, sizeof(uint64_t) * cap));
        FMM_TRY(cudaMalloc(&t.hkeys_alt, sizeof(uint64_t) * cap));
        FMM_TRY(cudaMalloc(&t.hperm, sizeof(int) * cap));
        FMM_TRY(cudaMalloc(&t.hperm_alt, sizeof(int) * cap));
        t.hcap = cap;
    }
    const int bins = 1 << (3 * Lc);
    if (bins != t.bins || !t.hist) {
        void* ptrs[] = {t.hist, t.hpre, t.binmax, t.split_idx, t.work};
        for (void* p : ptrs)
            if (p) cudaFree(p);
        t.hist = t.hpre = t.split_idx = nullptr;
        t.binmax = nullptr;
        t.work = nullptr;
        FMM_TRY(cudaMalloc(&t.work, sizeof(long long) * bins));
        FMM_TRY(cudaMemset(t.work, 0, sizeof(long long) * bins));
        FMM_TRY(cudaMalloc(&t.hist, sizeof(int) * bins));
        FMM_TRY(cudaMalloc(&t.hpre, sizeof(int) * (bins + 1)));
        FMM_TRY(cudaMalloc(&t.binmax, sizeof(double) * bins));
        FMM_TRY(cudaMalloc(&t.split_idx, sizeof(int) * 1024 + sizeof(uint64_t) * 1024));
        t.bins = bins;
    }
    t.Lc = Lc;
    return cudaSuccess;
}

// phase 1: local bounds of the home particles -> lohi[0..2] = min, lohi[3..5] = max (+-1e300 for an empty shard)
inline cudaError_t let_bounds(FmmWorkspace& w, const double* soa, int64_t ld, int64_t n, double* lohi, cudaStream_t st,
                              uint64_t& launches, std::string& err) {
    for (int c = 0; c < 3; ++c) { lohi[c] = 1e300; lohi[3 + c] = -1e300; }
    if (n <= 0) return cudaSuccess;
    FMM_TRY(fmm_reserve_particles(w, n, err));
    const int nb = 128;
    fmm_bounds_kernel<<<nb, 256, 0, st>>>(soa + (size_t)F_X * ld, soa + (size_t)(F_X + 1) * ld, soa + (size_t)(F_X + 2) * ld, n, w.bounds);
    ++launches;
    std::vector<double> hb(6 * nb);
    FMM_TRY(cudaMemcpyAsync(hb.data(), w.bounds, sizeof(double) * 6 * nb, cudaMemcpyDeviceToHost, st));
    FMM_TRY(cudaStreamSynchronize(st));
    for (int b = 0; b < nb; ++b)
        for (int c = 0; c < 3; ++c) {
            lohi[c] = std::min(lohi[c], hb[6 * b + 2 * c]);
            lohi[3 + c] = std::max(lohi[3 + c], -hb[6 * b + 2 * c + 1]);
        }
    return cudaSuccess;
}

// The root cube of fmm_sort from global bounds (same arithmetic, so the keys — and the tree — equal the one-GPU ones).
inline bool let_cube_from_bounds(const double* lohi, FmmRoot* rt) {
    for (int c = 0; c < 3; ++c)
        if (!(lohi[c] <= lohi[3 + c]) || !std::isfinite(lohi[c]) || !std::isfinite(lohi[3 + c])) return false;
    double side = std::max(std::max(lohi[3] - lohi[0], lohi[4] - lohi[1]), lohi[5] - lohi[2]);
    side = side > 0 ? side * (1.0 + 1e-9) : 1.0;
    rt->cx = 0.5 * (lohi[0] + lohi[3]);
    rt->cy = 0.5 * (lohi[1] + lohi[4]);
    rt->cz = 0.5 * (lohi[2] + lohi[5]);
    rt->side = side;
    return true;
}

// phase 2: keys + sort of the home particles in the global cube, level-Lc histogram (and per-bin sigma max)
inline cudaError_t let_keys(FmmWorkspace& w, FmmLet& t, const double* soa, int64_t ld, int64_t n, const FmmRoot& cube, int Lc,
                            bool want_binmax, cudaStream_t st, uint64_t& launches, std::string& err) {
    FMM_TRY(let_reserve_home(t, n, Lc, err));
    t.cube = cube;
    t.n_home = n;
    t.far_valid = false;
    const double x0 = cube.cx - 0.5 * cube.side, y0 = cube.cy - 0.5 * cube.side, z0 = cube.cz - 0.5 * cube.side;
    if (n > 0) {
        fmm_keys_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(soa + (size_t)F_X * ld, soa + (size_t)(F_X + 1) * ld,
                                                                    soa + (size_t)(F_X + 2) * ld, n, x0, y0, z0,
                                                                    2097152.0 / cube.side, t.hkeys, t.hperm);
        ++launches;
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, t.hkeys, t.hkeys_alt, t.hperm, t.hperm_alt, (int)n, 0, 63, st);
        FMM_TRY(fmm_cub(w, tb, err));
        tb = w.cub_bytes;
        FMM_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, t.hkeys, t.hkeys_alt, t.hperm, t.hperm_alt, (int)n, 0, 63, st));
        ++launches;
        std::swap(t.hkeys, t.hkeys_alt);
        std::swap(t.hperm, t.hperm_alt);
    }
    let_hist_kernel<<<(t.bins + 255) / 256, 256, 0, st>>>(t.hkeys, (int)n, Lc, t.bins, t.hist);
    ++launches;
    if (want_binmax) {
        let_binmax_kernel<<<(t.bins + 255) / 256, 256, 0, st>>>(t.hkeys, t.hperm, soa + (size_t)F_SIGMA * ld, (int)n, Lc, t.bins, t.binmax);
        ++launches;
    }
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

// Units of the global top tree in Morton order: a level-Lc bin whose ancestors are all split, or a coarser cell that is a
// leaf of the global tree (<= ncrit particles over all ranks).  cnt[l] = per-cell counts of level l (Morton-indexed).
inline void let_units(const std::vector<std::vector<int>>& cnt, int Lc, int ncrit, int level, uint64_t q,
                      std::vector<std::pair<uint64_t, int64_t>>& units /* (first key, count) */) {
    const int c = cnt[level][q];
    if (c == 0) return;
    if (level == Lc || c <= ncrit) {
        units.emplace_back(q << (3 * (FMM_MAXLEVEL - level)), (int64_t)c);
        return;
    }
    for (int o = 0; o < 8; ++o) let_units(cnt, Lc, ncrit, level + 1, (q << 3) | (uint64_t)o, units);
}

// The cut of the Morton curve — pure host arithmetic on the all-reduced histogram (and, optionally, the all-reduced work counts)
// of the level-Lc bins, so every rank derives the SAME splitters; exported as vpmb200_let_cut for the CPU tests.
// splitters: nparts + 1 key bounds (rank k owns keys in [splitters[k], splitters[k + 1])).
inline void let_cut(const int* hist, const long long* work, int Lc, int ncrit, int nparts, std::vector<uint64_t>& splitters,
                    std::vector<int>* prefix_out = nullptr) {
    const int bins = 1 << (3 * Lc);
    std::vector<std::vector<int>> cnt(Lc + 1);
    cnt[Lc].assign(hist, hist + bins);
    for (int l = Lc - 1; l >= 0; --l) {
        cnt[l].assign((size_t)1 << (3 * l), 0);
        for (size_t q = 0; q < cnt[l + 1].size(); ++q) cnt[l][q >> 3] += cnt[l + 1][q];
    }
    const int64_t ntot = cnt[0][0];
    if (prefix_out) {
        prefix_out->assign(bins + 1, 0);
        for (int b = 0; b < bins; ++b) (*prefix_out)[b + 1] = (*prefix_out)[b] + cnt[Lc][b];
    }
    std::vector<std::pair<uint64_t, int64_t>> units;
    if (ntot > 0) let_units(cnt, Lc, ncrit, 0, 0, units);
    // Weight of a unit: the interaction work counted in its bins during the previous evaluation (all-reduced by the caller)
    // blended with its particle count — equal particle counts are not equal work (the near-field cost follows leaf occupancy).
    // Without counted work (first evaluation) the cut is by count.
    std::vector<double> wgt(units.size());
    double wtot = 0.0;
    {
        double per_particle = 0.0;
        std::vector<double> wpre;
        if (work && ntot > 0) {
            wpre.assign(bins + 1, 0.0);
            for (int bq = 0; bq < bins; ++bq) wpre[bq + 1] = wpre[bq] + (double)work[bq];
            per_particle = wpre[bins] / (double)ntot;
        }
        for (size_t u = 0; u < units.size(); ++u) {
            double w = (double)units[u].second;
            if (per_particle > 0.0) {
                // bins from the unit's first bin up to the next unit's first bin (the bins in between hold no particles)
                const uint64_t b0 = units[u].first >> (3 * (FMM_MAXLEVEL - Lc));
                const uint64_t b1 = u + 1 < units.size() ? units[u + 1].first >> (3 * (FMM_MAXLEVEL - Lc)) : (uint64_t)bins;
                w = (wpre[b1] - wpre[b0]) + 0.25 * per_particle * (double)units[u].second;
            }
            wgt[u] = w;
            wtot += w;
        }
    }
    // rank k takes the units whose cumulative weight (at the unit's START) falls in [k, k + 1) * wtot / nparts
    splitters.assign(nparts + 1, ~0ull >> 1);
    splitters[0] = 0;
    {
        double cum = 0.0;
        int k = 1;
        for (size_t u = 0; u < units.size(); ++u) {
            while (k < nparts && cum >= wtot * k / nparts) splitters[k++] = units[u].first;
            cum += wgt[u];
        }
        // ranks left without a unit get empty ranges at the end of the curve
    }
    splitters[nparts] = 1ull << 63;
    for (int k = 1; k < nparts; ++k)
        if (splitters[k] == (~0ull >> 1)) splitters[k] = 1ull << 63;
}

// phase 3 (after the caller all-reduced t.hist in place): splitters, send counts, prefix sums for the forced top splits
inline cudaError_t let_partition(FmmLet& t, int nparts, int part, int ncrit, bool use_work, int64_t* send_counts, cudaStream_t st,
                                 uint64_t& launches, std::string& err) {
    if (nparts > 1024) { err = "LET: more than 1024 ranks"; return cudaErrorInvalidValue; }
    const int bins = t.bins, Lc = t.Lc;
    std::vector<int> hh(bins);
    FMM_TRY(cudaMemcpyAsync(hh.data(), t.hist, sizeof(int) * bins, cudaMemcpyDeviceToHost, st));
    std::vector<long long> hw;
    if (use_work && t.work) {
        hw.resize(bins);
        FMM_TRY(cudaMemcpyAsync(hw.data(), t.work, sizeof(long long) * bins, cudaMemcpyDeviceToHost, st));
    }
    FMM_TRY(cudaStreamSynchronize(st));
    std::vector<int> pre;
    let_cut(hh.data(), hw.empty() ? nullptr : hw.data(), Lc, ncrit, nparts, t.splitters, &pre);
    FMM_TRY(cudaMemcpyAsync(t.hpre, pre.data(), sizeof(int) * (bins + 1), cudaMemcpyHostToDevice, st));
    t.nparts = nparts;
    t.part = part;
    // send counts: where the splitters fall in the sorted home keys
    uint64_t* dsplit = reinterpret_cast<uint64_t*>(t.split_idx + 1024);
    FMM_TRY(cudaMemcpyAsync(dsplit, t.splitters.data(), sizeof(uint64_t) * (nparts + 1), cudaMemcpyHostToDevice, st));
    let_find_splits_kernel<<<(nparts + 1 + 127) / 128, 128, 0, st>>>(t.hkeys, (int)t.n_home, dsplit, nparts + 1, t.split_idx);
    ++launches;
    std::vector<int> idx(nparts + 1);
    FMM_TRY(cudaMemcpyAsync(idx.data(), t.split_idx, sizeof(int) * (nparts + 1), cudaMemcpyDeviceToHost, st));
    FMM_TRY(cudaStreamSynchronize(st));
    t.send_counts.resize(nparts);
    for (int k = 0; k < nparts; ++k) {
        t.send_counts[k] = idx[k + 1] - idx[k];
        if (send_counts) send_counts[k] = t.send_counts[k];
    }
    return cudaSuccess;
}

inline int let_nm(int p) {
    switch (p) {
    case 2: return FmmOps<2>::NM;
    case 3: return FmmOps<3>::NM;
    case 4: return FmmOps<4>::NM;
    case 5: return FmmOps<5>::NM;
    default: return FmmOps<6>::NM;
    }
}

inline cudaError_t let_upward(FmmWorkspace& w, int p, const std::vector<int>& lvl, cudaStream_t st, uint64_t& launches) {
    switch (p) {
    case 2: return FmmPasses<2>::upward(w, lvl, st, launches);
    case 3: return FmmPasses<3>::upward(w, lvl, st, launches);
    case 4: return FmmPasses<4>::upward(w, lvl, st, launches);
    case 5: return FmmPasses<5>::upward(w, lvl, st, launches);
    case 6: return FmmPasses<6>::upward(w, lvl, st, launches);
    default: return cudaErrorInvalidValue;
    }
}

// phases 4 + 5: sort the received rows, build the owner's part of the global tree, upward pass.
// n_all: particles of ALL ranks (capacity of the combined record array).  reuse: positions and strengths are those of the
// previous evaluation (DynamicSFS domain-filter evaluation): keep order, tree, lists and local expansions, refresh the records.
inline cudaError_t let_build(FmmWorkspace& w, FmmLet& t, const double* rows, int64_t n_own, int64_t n_all, int ncrit,
                             double nzs_factor, int p, bool reuse, cudaStream_t st, uint64_t& launches, std::string& err) {
    t.rows = rows;
    const unsigned nbk = (unsigned)((n_own + 255) / 256);
    if (reuse) {
        if (!t.far_valid || n_own != t.n_own || p != t.P) { err = "LET: nothing to reuse"; return cudaErrorInvalidValue; }
        if (n_own > 0) {
            let_gather_rows_kernel<<<nbk, 256, 0, st>>>(rows, n_own, w.perm, w.sx, w.sy, w.sz, w.rec);
            ++launches;
        }
        return cudaGetLastError();
    }
    t.far_valid = false;
    t.n_own = n_own;
    t.n_all = n_all;
    t.P = p;
    t.ncells_own = 0;
    w.ncells = w.nleaves = w.nlevels = 0;
    w.leaf_lo = w.leaf_hi = 0;
    t.lvl.assign(1, 0);
    FMM_TRY(fmm_reserve_particles(w, std::max<int64_t>(n_all, 1), err));
    if (n_own <= 0) return cudaSuccess;
    const double x0 = t.cube.cx - 0.5 * t.cube.side, y0 = t.cube.cy - 0.5 * t.cube.side, z0 = t.cube.cz - 0.5 * t.cube.side;
    let_keys_rows_kernel<<<nbk, 256, 0, st>>>(rows, n_own, x0, y0, z0, 2097152.0 / t.cube.side, w.keys, w.perm);
    ++launches;
    {
        size_t tb = 0;
        cub::DeviceRadixSort::SortPairs(nullptr, tb, w.keys, w.keys_alt, w.perm, w.perm_alt, (int)n_own, 0, 63, st);
        FMM_TRY(fmm_cub(w, tb, err));
        tb = w.cub_bytes;
        FMM_TRY(cub::DeviceRadixSort::SortPairs(w.cub_tmp, tb, w.keys, w.keys_alt, w.perm, w.perm_alt, (int)n_own, 0, 63, st));
        ++launches;
        std::swap(w.keys, w.keys_alt);
        std::swap(w.perm, w.perm_alt);
    }
    let_gather_rows_kernel<<<nbk, 256, 0, st>>>(rows, n_own, w.perm, w.sx, w.sy, w.sz, w.rec);
    ++launches;
    {
        cudaError_t e1 = fmm_tree(w, n_own, ncrit, t.cube, t.lvl, st, launches, err, t.hpre, t.Lc);
        if (e1 != cudaSuccess) return e1;
    }
    t.ncells_own = w.ncells;
    if (nzs_factor > 0.0) {
        fmm_smax(w, w.cells, w.ncells, t.lvl, st, launches);
        let_top_smax_kernel<<<(w.ncells + 127) / 128, 128, 0, st>>>(w.cells, w.ncells, w.keys, t.Lc, t.binmax);
        ++launches;
    }
    let_stamp_leaves_kernel<<<(w.ncells + 255) / 256, 256, 0, st>>>(w.cells, w.ncells, w.leaf_pos);   // (for the halo mode)
    ++launches;
    fmm_tic(w, 2, st);
    cudaError_t e2 = let_upward(w, p, t.lvl, st, launches);
    fmm_toc(w, 2, st);
    if (e2 != cudaSuccess) { err = std::string("LET upward pass: ") + cudaGetErrorString(e2); return e2; }
    return cudaSuccess;
}

// phase 6a: the other ranks' tree skeletons and multipoles, appended behind the own ones.
// cells_recv / M_recv: all-gathered buffers, rank q's block at q * slot_cells cells / q * slot_cells * 3 NM doubles.
inline cudaError_t let_attach_tree(FmmWorkspace& w, FmmLet& t, const FmmCell* cells_recv, const double* M_recv, int64_t slot_cells,
                                   const int64_t* ncells, const int64_t* nparticles, cudaStream_t st, uint64_t& launches,
                                   std::string& err) {
    const int G = t.nparts, nm3 = 3 * let_nm(t.P);
    int64_t total = 0, ntot = 0;
    for (int q = 0; q < G; ++q) { total += ncells[q]; ntot += nparticles[q]; }
    if (ncells[t.part] != t.ncells_own || nparticles[t.part] != t.n_own) { err = "LET: own block sizes do not match"; return cudaErrorInvalidValue; }
    if (ntot > w.cap_n) { err = "LET: record array too small for the gathered particles"; return cudaErrorInvalidValue; }
    if (total > t.cap_cells_all || !t.cells_all) {
        if (t.cells_all) cudaFree(t.cells_all);
        t.cells_all = nullptr;
        t.cap_cells_all = 0;
        const int64_t cap = total + total / 4 + 64;
        FMM_TRY(cudaMalloc(&t.cells_all, sizeof(FmmCell) * cap));
        t.cap_cells_all = cap;
    }
    FMM_TRY(fmm_grow(t.M_all, t.cap_M_all, (size_t)std::max<int64_t>(total, 1) * nm3, err));
    t.cell_off.assign(G, 0);
    t.part_off.assign(G, 0);
    t.ncells_of.assign(ncells, ncells + G);
    if (t.ncells_own > 0) {
        FMM_TRY(cudaMemcpyAsync(t.cells_all, w.cells, sizeof(FmmCell) * t.ncells_own, cudaMemcpyDeviceToDevice, st));
        FMM_TRY(cudaMemcpyAsync(t.M_all, w.M, sizeof(double) * (size_t)t.ncells_own * nm3, cudaMemcpyDeviceToDevice, st));
    }
    int coff = t.ncells_own;
    int64_t poff = t.n_own;
    for (int q = 0; q < G; ++q) {
        if (q == t.part) continue;
        t.cell_off[q] = coff;
        t.part_off[q] = poff;
        const int nc = (int)ncells[q];
        if (nc > 0) {
            let_attach_cells_kernel<<<(nc + 255) / 256, 256, 0, st>>>(cells_recv + (size_t)q * slot_cells, nc, t.cells_all + coff, coff, (int)poff);
            ++launches;
            FMM_TRY(cudaMemcpyAsync(t.M_all + (size_t)coff * nm3, M_recv + (size_t)q * slot_cells * nm3, sizeof(double) * (size_t)nc * nm3,
                                    cudaMemcpyDeviceToDevice, st));
            let_leaf_counts_kernel<<<(nc + 255) / 256, 256, 0, st>>>(t.cells_all, coff, coff + nc, w.count_at);
            ++launches;
        }
        coff += nc;
        poff += nparticles[q];
    }
    t.ncells_all = coff;
    t.halo_mode = false;
    w.halo = FmmHalo();
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

// phase 6b: the other ranks' source records (UJ records, or E_str records for the second pass) behind the own ones in w.rec.
// rec_recv: all-gathered, rank q's block at q * slot_n records.
inline cudaError_t let_attach_records(FmmWorkspace& w, FmmLet& t, const double* rec_recv, int64_t slot_n, const int64_t* nparticles,
                                      cudaStream_t st, std::string& err) {
    for (int q = 0; q < t.nparts; ++q) {
        if (q == t.part || nparticles[q] <= 0) continue;
        FMM_TRY(cudaMemcpyAsync(w.rec + (size_t)t.part_off[q] * REC_REALS, rec_recv + (size_t)q * slot_n * REC_REALS,
                                sizeof(double) * (size_t)nparticles[q] * REC_REALS, cudaMemcpyDeviceToDevice, st));
    }
    return cudaSuccess;
}

// phase 7: U, J of the owner's particles; out: n_own rows of 12 (U, J) in the order the rows arrived.
// stage 0: everything; stage 1: interaction lists + far field (traversal, M2L, L2L — needs the skeletons and multipoles only,
// so the caller can keep the source records in flight meanwhile); stage 2: L2P + near field + output rows.
inline cudaError_t let_evaluate(FmmWorkspace& w, FmmLet& t, double theta, double nzs_factor, int kernel, int block,
                                const double* gh_table, double* out, bool reuse, int stage, cudaStream_t st, uint64_t& launches,
                                std::string& err) {
    if (t.n_own <= 0) return cudaSuccess;
    if (stage != 2 && stage != 6 && !reuse) {
        std::vector<uint64_t> seeds;
        seeds.push_back(0);                                      // (own root, own root)
        for (int q = 0; q < t.nparts; ++q)                       // (own root, root of rank q's tree); empty ranks have no tree
            if (q != t.part && t.ncells_of[q] > 0) seeds.push_back((uint64_t)(unsigned int)t.cell_off[q]);
        cudaError_t e0 = fmm_lists(w, t.cells_all, t.ncells_own, theta, nzs_factor, nullptr, seeds.data(), (int)seeds.size(), 1, st, launches, err);
        if (e0 != cudaSuccess) return e0;
        if (t.work) {   // (zeroed by the caller before the evaluation, so ranks without particles contribute zeros)
            let_count_work_kernel<<<(t.ncells_own + 255) / 256, 256, 0, st>>>(t.cells_all, t.ncells_own, w.p2p_off, w.runs, w.m2l_off, w.keys,
                                                                            t.Lc, t.work);
            ++launches;
        }
    }
    if (stage == 5) return cudaSuccess;                          // lists only (the halo is planned from them)
    w.cells_eval = t.cells_all;
    w.M_eval = t.halo_mode ? w.M : t.M_all;                      // halo mode: own multipoles in place, the others' in w.halo.M2
    cudaError_t e1 = fmm_evaluate(w, t.P, kernel, block, gh_table, t.lvl, st, launches, reuse, /*skip_upward=*/true,
                                  stage == 6 ? 1 : stage);
    w.cells_eval = nullptr;
    w.M_eval = nullptr;
    if (e1 != cudaSuccess) { err = std::string("LET evaluate: ") + cudaGetErrorString(e1); return e1; }
    if (stage == 1 || stage == 6) return cudaSuccess;
    t.far_valid = true;
    let_out_rows_kernel<<<(unsigned)((t.n_own + 255) / 256), 256, 0, st>>>(w.sU, 3, w.sJ, 9, w.lds, t.n_own, w.perm, out);
    ++launches;
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

// E_str records of the owner's particles from the TOTAL J (w.sJ after let_evaluate; `accumulated` J would need the home
// rank's previous J — UJ with sfs = true always resets, as every SFS scheme calls it) into w.rec[0, n_own).
inline cudaError_t let_estr_records(FmmWorkspace& w, FmmLet& t, int transposed, double zeta_norm, cudaStream_t st, uint64_t& launches) {
    if (t.n_own <= 0) return cudaSuccess;
    let_estr_records_kernel<<<(unsigned)((t.n_own + 255) / 256), 256, 0, st>>>(t.rows, t.n_own, w.perm, w.sJ, w.lds, transposed, zeta_norm, w.rec);
    ++launches;
    return cudaGetLastError();
}

// near-field E_str of the owner's particles over the same leaf pairs; out: n_own rows of 3 in arrival order
inline cudaError_t let_estr_evaluate(FmmWorkspace& w, FmmLet& t, int kernel, int block, int transposed, const double* z_table,
                                     double* out, cudaStream_t st, uint64_t& launches, std::string& err) {
    if (t.n_own <= 0) return cudaSuccess;
    w.cells_eval = t.cells_all;
    cudaError_t e1 = fmm_estr(w, kernel, block, transposed, z_table, st, launches);
    w.cells_eval = nullptr;
    if (e1 != cudaSuccess) { err = std::string("LET E_str: ") + cudaGetErrorString(e1); return e1; }
    let_out_rows_kernel<<<(unsigned)((t.n_own + 255) / 256), 256, 0, st>>>(w.sE, 3, w.sE, 0, w.lds, t.n_own, w.perm, out);
    ++launches;
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

// ---- demand-driven halo: host side ---------------------------------------------------------------------------------------
template <typename T>
inline cudaError_t let_grow_arr(T*& p, int64_t& cap, int64_t need, std::string& err) {
    if (need <= cap && p) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const int64_t c = need + need / 4 + 1024;
    FMM_TRY(cudaMalloc(&p, sizeof(T) * c));
    cap = c;
    return cudaSuccess;
}

// phase 6a (halo mode): the other ranks' SKELETONS only.  cells_recv: all-gathered, rank q's block at q * slot_cells.
inline cudaError_t let_attach_skeleton(FmmWorkspace& w, FmmLet& t, const FmmCell* cells_recv, int64_t slot_cells, const int64_t* ncells,
                                       const int64_t* nparticles, const int64_t* nleaves, cudaStream_t st, uint64_t& launches,
                                       std::string& err) {
    const int G = t.nparts;
    int64_t total = 0, nl = 0;
    for (int q = 0; q < G; ++q) {
        total += ncells[q];
        if (q != t.part) nl += nleaves[q];
    }
    if (ncells[t.part] != t.ncells_own || nparticles[t.part] != t.n_own) { err = "LET: own block sizes do not match"; return cudaErrorInvalidValue; }
    if (total > t.cap_cells_all || !t.cells_all) {
        if (t.cells_all) cudaFree(t.cells_all);
        t.cells_all = nullptr;
        t.cap_cells_all = 0;
        const int64_t cap = total + total / 4 + 64;
        FMM_TRY(cudaMalloc(&t.cells_all, sizeof(FmmCell) * cap));
        t.cap_cells_all = cap;
    }
    {   // per-cell and per-remote-leaf bookkeeping (+1: the scans read one element past the end for the totals)
        int64_t c1 = t.cap_hc, c2 = t.cap_hc, c3 = t.cap_hc;
        FMM_TRY(let_grow_arr(t.mneed, c1, total + 1, err));
        FMM_TRY(let_grow_arr(t.mslot, c2, total + 1, err));
        FMM_TRY(let_grow_arr(t.req_cells, c3, total + 1, err));
        t.cap_hc = std::min(c1, std::min(c2, c3));
        int64_t l1 = t.cap_hl, l2 = t.cap_hl, l3 = t.cap_hl, l4 = t.cap_hl, l5 = t.cap_hl, l6 = t.cap_hl;
        FMM_TRY(let_grow_arr(t.rleaf, l1, nl + 1, err));
        FMM_TRY(let_grow_arr(t.pneed, l2, nl + 1, err));
        FMM_TRY(let_grow_arr(t.pslot, l3, nl + 1, err));
        FMM_TRY(let_grow_arr(t.pcnt, l4, nl + 1, err));
        FMM_TRY(let_grow_arr(t.poff, l5, nl + 1, err));
        FMM_TRY(let_grow_arr(t.req_leaf, l6, nl + 1, err));
        t.cap_hl = std::min(std::min(l1, l2), std::min(std::min(l3, l4), std::min(l5, l6)));
    }
    if (t.n_own + nl > w.count_at_cap) {   // count_at is indexed by n_own + remote leaf ordinal in this mode
        if (w.count_at) cudaFree(w.count_at);
        w.count_at = nullptr;
        w.count_at_cap = 0;
        const int64_t cap = t.n_own + nl + (t.n_own + nl) / 4 + 1024;
        FMM_TRY(cudaMalloc(&w.count_at, sizeof(int) * cap));
        w.count_at_cap = cap;
    }
    t.cell_off.assign(G, 0);
    t.part_off.assign(G, 0);
    t.leaf_off.assign(G, 0);
    t.ncells_of.assign(ncells, ncells + G);
    t.nleaves_of.assign(nleaves, nleaves + G);
    if (t.ncells_own > 0) FMM_TRY(cudaMemcpyAsync(t.cells_all, w.cells, sizeof(FmmCell) * t.ncells_own, cudaMemcpyDeviceToDevice, st));
    int coff = t.ncells_own, loff = 0;
    for (int q = 0; q < G; ++q) {
        if (q == t.part) continue;
        t.cell_off[q] = coff;
        t.leaf_off[q] = loff;
        const int nc = (int)ncells[q];
        if (nc > 0) {
            let_attach_skeleton_kernel<<<(nc + 255) / 256, 256, 0, st>>>(cells_recv + (size_t)q * slot_cells, nc, t.cells_all + coff, coff, loff,
                                                                       (int)t.n_own, t.rleaf, w.count_at);
            ++launches;
        }
        coff += nc;
        loff += (int)nleaves[q];
    }
    t.ncells_all = coff;
    t.nleaves_remote = loff;
    t.halo_mode = true;
    w.halo = FmmHalo();
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

// After the interaction lists exist (let_evaluate stage 5): which remote multipoles and which remote leaves' records does this
// rank need?  counts3[3 q + 0 / 1 / 2] = cells / leaves / records requested from rank q (zeros for the rank itself); the request
// arrays t.req_cells / t.req_leaf are grouped by owner in rank order; the P2P runs of remote leaves are rewritten to point into
// the halo record buffer (in request order).
inline cudaError_t let_halo_plan(FmmWorkspace& w, FmmLet& t, int64_t* counts3, cudaStream_t st, uint64_t& launches, std::string& err) {
    const int G = t.nparts;
    for (int q = 0; q < 3 * G; ++q) counts3[q] = 0;
    if (!t.halo_mode) { err = "LET: halo plan without a skeleton attach"; return cudaErrorInvalidValue; }
    const int nc = t.ncells_all, nl = t.nleaves_remote;
    FMM_TRY(cudaMemsetAsync(t.mneed, 0, sizeof(int) * (nc + 1), st));
    FMM_TRY(cudaMemsetAsync(t.pneed, 0, sizeof(int) * (nl + 1), st));
    FMM_TRY(cudaMemsetAsync(t.pcnt, 0, sizeof(int) * (nl + 1), st));
    if (t.n_own > 0) {
        if (w.n_m2l > 0) let_mark_m2l_kernel<<<(w.n_m2l + 255) / 256, 256, 0, st>>>(w.m2l_sorted, w.n_m2l, t.ncells_own, t.mneed);
        if (w.n_p2p > 0) let_mark_p2p_kernel<<<(w.n_p2p + 255) / 256, 256, 0, st>>>(w.p2p_sorted, w.n_p2p, (int)t.n_own, t.rleaf, t.pneed, t.pcnt);
        launches += 2;
    }
    FMM_CUB(cub::DeviceScan::ExclusiveSum(tmp, tb, t.mneed, t.mslot, nc + 1, st));
    FMM_CUB(cub::DeviceScan::ExclusiveSum(tmp, tb, t.pneed, t.pslot, nl + 1, st));
    FMM_CUB(cub::DeviceScan::ExclusiveSum(tmp, tb, t.pcnt, t.poff, nl + 1, st));
    launches += 3;
    // per-owner totals: differences of the scans at the block boundaries
    std::vector<int> hb(6 * G, 0);
    for (int q = 0; q < G; ++q) {
        if (q == t.part) continue;
        const int c0 = t.cell_off[q], c1 = c0 + (int)t.ncells_of[q], l0 = t.leaf_off[q], l1 = l0 + (int)t.nleaves_of[q];
        FMM_TRY(cudaMemcpyAsync(&hb[6 * q + 0], t.mslot + c0, sizeof(int), cudaMemcpyDeviceToHost, st));
        FMM_TRY(cudaMemcpyAsync(&hb[6 * q + 1], t.mslot + c1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FMM_TRY(cudaMemcpyAsync(&hb[6 * q + 2], t.pslot + l0, sizeof(int), cudaMemcpyDeviceToHost, st));
        FMM_TRY(cudaMemcpyAsync(&hb[6 * q + 3], t.pslot + l1, sizeof(int), cudaMemcpyDeviceToHost, st));
        FMM_TRY(cudaMemcpyAsync(&hb[6 * q + 4], t.poff + l0, sizeof(int), cudaMemcpyDeviceToHost, st));
        FMM_TRY(cudaMemcpyAsync(&hb[6 * q + 5], t.poff + l1, sizeof(int), cudaMemcpyDeviceToHost, st));
        if (t.ncells_of[q] > 0) {
            let_compact_cells_kernel<<<((int)t.ncells_of[q] + 255) / 256, 256, 0, st>>>(t.mneed, t.mslot, c0, c1, t.req_cells);
            ++launches;
        }
    }
    if (nl > 0) {
        let_compact_leaves_kernel<<<(nl + 255) / 256, 256, 0, st>>>(t.pneed, t.pslot, t.rleaf, nl, t.req_leaf);
        ++launches;
    }
    if (t.n_own > 0 && w.n_p2p > 0) {
        let_halo_runs_kernel<<<(w.n_p2p + 255) / 256, 256, 0, st>>>(w.p2p_sorted, w.n_p2p, (int)t.n_own, t.rleaf, t.poff, w.runs);
        ++launches;
    }
    FMM_TRY(cudaStreamSynchronize(st));
    for (int q = 0; q < G; ++q) {
        counts3[3 * q + 0] = hb[6 * q + 1] - hb[6 * q + 0];
        counts3[3 * q + 1] = hb[6 * q + 3] - hb[6 * q + 2];
        counts3[3 * q + 2] = hb[6 * q + 5] - hb[6 * q + 4];
    }
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

// Owner side: the multipoles of the requested cells (ids local to this rank's tree; M_out may be NULL when only records are
// re-served) and the records of the requested leaves, back to back in request order.
inline cudaError_t let_halo_serve(FmmWorkspace& w, FmmLet& t, const int* req_cells, int64_t ncell, const int2* req_leaf, int64_t nleaf,
                                  double* M_out, double* rec_out, cudaStream_t st, uint64_t& launches, std::string& err) {
    const int nm3 = 3 * let_nm(t.P);
    if (M_out && ncell > 0) {
        const int64_t tot = ncell * nm3;
        let_serve_M_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, st>>>(req_cells, (int)ncell, nm3, w.M, M_out);
        ++launches;
    }
    if (rec_out && nleaf > 0) {
        FMM_TRY(let_grow_arr(t.serve_off, t.cap_serve, 2 * (nleaf + 1), err));
        int* cnt = t.serve_off + (nleaf + 1);
        let_serve_counts_kernel<<<(unsigned)((nleaf + 255) / 256), 256, 0, st>>>(req_leaf, (int)nleaf, cnt);
        FMM_CUB(cub::DeviceScan::ExclusiveSum(tmp, tb, cnt, t.serve_off, (int)nleaf, st));
        let_serve_rec_kernel<<<(unsigned)((nleaf + 3) / 4), 128, 0, st>>>(req_leaf, t.serve_off, (int)nleaf, w.rec, rec_out);
        launches += 3;
    }
    FMM_TRY(cudaGetLastError());
    return cudaSuccess;
}

}  // namespace vpm
