// fmm.cuh — K3: GPU fast multipole evaluation of U and J (FLOWVPM's UJ_fmm, selected by `vpm_UJ = vpm.UJ_fmm` with
// settings `vpm.FMM(; p=4, ncrit=50, theta=0.4, nonzero_sigma=false)`, /root/reference/src/FLOWUnsteady_simulation.jl:38,43;
// the reference delegates to ExaFMM (C++/OpenMP, README.md:115-116) or FastMultipole.jl, neither of which is in
// the tree — SURVEY.md §0.4).  What is matched is the ALGORITHM CLASS and its parameters, not bits:
//   * adaptive octree over Morton-sorted particles, leaves hold <= ncrit particles;
//   * well-separated test (R_i + R_j) < theta |c_i - c_j| with R the cell's half side (ExaFMM's convention);
//   * far field: singular (1/r) kernel of the vector potential psi = (1/4 pi) sum Gamma/r  (`nonzero_sigma = false`),
//     Cartesian Taylor expansions with multipoles to order p-1 and locals to order p+1, U = curl psi and J = grad U taken
//     analytically from the local expansion (the reference obtains J by complex-step differentiation, rvpm.md:370);
//   * near field: the regularised pair kernel of uj_direct.cuh (same device functions) over leaf pairs;
//   * E_str (`sfs = true`): a second near-field pass over the same leaf pairs (Estr_fmm evaluates the SFS term in the
//     near field only).
// Everything runs on the device; the host only reads back a few counters per tree level / traversal sweep.
// Deterministic: child order, pair lists (sorted by (target, source)) and all reductions have a fixed order.
#pragma once

#include <cub/cub.cuh>

#include "estr_direct.cuh"
#include "fmm_ops.inc"
#include "uj_direct.cuh"

namespace vpm {

constexpr int FMM_MAXLEVEL = 21;      // 3 x 21 = 63-bit Morton keys
constexpr int FMM_MAX_NCRIT = 256;

struct FmmCell {
    int start, count;    // particle range in Morton order
    int parent, child0;  // first child (children are contiguous), -1 for none
    int nchild, level;
    double cx, cy, cz, R;  // centre and half side
    double smax;           // largest core size sigma among the cell's particles (regularisation-aware acceptance)
    double pad_;
};

// Local essential tree, demand-driven (fmm_let.cuh): sources beyond the rank's own block live in separate RECEIVED buffers.
// Near field: a run whose first record index is >= rec_split reads rec2[index - rec_split]; M2L: a source cell id >= cell_split
// reads the multipole M2[mslot[id]].  The defaults switch both off (single GPU, all-gather LET).
struct FmmHalo {
    int rec_split = 0x7fffffff;
    const double* rec2 = nullptr;
    int cell_split = 0x7fffffff;
    const double* M2 = nullptr;
    const int* mslot = nullptr;
};

// ---- bounding box -----------------------------------------------------------------------------------------------
__global__ void fmm_bounds_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                  int64_t n, double* __restrict__ out /* gridDim.x * 6 */) {
    __shared__ double red[6][8];
    const double big = 1.0e300;
    double v[6] = {big, big, big, big, big, big};  // min x, -max x, min y, -max y, min z, -max z
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double a = x[i], b = y[i], c = z[i];
        v[0] = fmin(v[0], a); v[1] = fmin(v[1], -a);
        v[2] = fmin(v[2], b); v[3] = fmin(v[3], -b);
        v[4] = fmin(v[4], c); v[5] = fmin(v[5], -c);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] = fmin(v[c], __shfl_xor_sync(0xffffffffu, v[c], o));
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 6; ++c) red[c][threadIdx.x >> 5] = v[c];
    __syncthreads();
    if (threadIdx.x < 6) {
        double m = red[threadIdx.x][0];
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmin(m, red[threadIdx.x][k]);
        out[blockIdx.x * 6 + threadIdx.x] = m;
    }
}

// ---- Morton keys ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t spread3(uint64_t v) {  // 21 bits -> every third bit
    v &= 0x1fffffull;
    v = (v | (v << 32)) & 0x1f00000000ffffull;
    v = (v | (v << 16)) & 0x1f0000ff0000ffull;
    v = (v | (v << 8)) & 0x100f00f00f00f00full;
    v = (v | (v << 4)) & 0x10c30c30c30c30c3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}

__global__ void fmm_keys_kernel(const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                                int64_t n, double x0, double y0, double z0, double inv_cell /* 2^21 / side */,
                                uint64_t* __restrict__ keys, int* __restrict__ perm) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double lim = 2097151.0;
    uint64_t ix = (uint64_t)fmin(fmax((x[i] - x0) * inv_cell, 0.0), lim);
    uint64_t iy = (uint64_t)fmin(fmax((y[i] - y0) * inv_cell, 0.0), lim);
    uint64_t iz = (uint64_t)fmin(fmax((z[i] - z0) * inv_cell, 0.0), lim);
    keys[i] = (spread3(ix) << 2) | (spread3(iy) << 1) | spread3(iz);
    perm[i] = (int)i;
}

// ---- gather into Morton order: target positions + UJ source records (common.cuh layout, no tile headers) -------------
__global__ void fmm_gather_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, const int* __restrict__ perm,
                                  double* __restrict__ sx, double* __restrict__ sy, double* __restrict__ sz,
                                  double* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = perm[i];
    double x = soa[(size_t)(F_X + 0) * ld + p], y = soa[(size_t)(F_X + 1) * ld + p], z = soa[(size_t)(F_X + 2) * ld + p];
    double gx = soa[(size_t)(F_GAMMA + 0) * ld + p], gy = soa[(size_t)(F_GAMMA + 1) * ld + p], gz = soa[(size_t)(F_GAMMA + 2) * ld + p];
    double sg = soa[(size_t)F_SIGMA * ld + p];
    double si = 1.0 / sg, si2 = si * si, si3 = si2 * si;
    sx[i] = x; sy[i] = y; sz[i] = z;
    double2* r = reinterpret_cast<double2*>(rec + (size_t)i * REC_REALS);
    r[0] = make_double2(x, y);
    r[1] = make_double2(z, -CONST4 * gx);
    r[2] = make_double2(-CONST4 * gy, -CONST4 * gz);
    r[3] = make_double2(VPM_GT_TFAR * (sg * sg), si3);
    r[4] = make_double2(si3 * si2, si2);
}

// E_str records in Morton order (estr_direct.cuh layout).  J is the field's CURRENT J (state rows, particle order), which
// is also copied into Morton order (sJ) for the targets of the second pass.
__global__ void fmm_gather_estr_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, const int* __restrict__ perm,
                                       double* __restrict__ sJ, int64_t ldj, int transposed, double zeta_norm,
                                       double* __restrict__ rec) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = perm[i];
    double x = soa[(size_t)(F_X + 0) * ld + p], y = soa[(size_t)(F_X + 1) * ld + p], z = soa[(size_t)(F_X + 2) * ld + p];
    double g0 = soa[(size_t)(F_GAMMA + 0) * ld + p], g1 = soa[(size_t)(F_GAMMA + 1) * ld + p], g2 = soa[(size_t)(F_GAMMA + 2) * ld + p];
    double sg = soa[(size_t)F_SIGMA * ld + p];
    double Jq[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        Jq[c] = soa[(size_t)(F_J + c) * ld + p];
        sJ[(size_t)c * ldj + i] = Jq[c];
    }
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

// ---- tree build: one level at a time -------------------------------------------------------------------------------
__device__ __forceinline__ int lower_bound_key(const uint64_t* __restrict__ keys, int lo, int hi, uint64_t v) {
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Is this cell split?  Single GPU: more than ncrit particles.  Local essential tree (fmm_let.cuh): the cells above level
// `Lc` belong to the GLOBAL tree's top, which is split by the GLOBAL particle count — `hpre` is the exclusive prefix sum of
// the all-reduced level-Lc histogram in Morton order, so a cell at level l with Morton prefix q holds
// hpre[(q + 1) << 3 (Lc - l)] - hpre[q << 3 (Lc - l)] particles over all ranks — so that every rank's partial top tree has
// exactly the structure of the one-GPU tree.
// Sparse-leaf refinement (`rec` != nullptr): a cell that would be a leaf (<= ncrit particles) is split anyway while its half
// side exceeds FMM_LEAF_SIGMAS core sizes of its own particles.  A few stray particles in a large empty cell next to a dense
// wake otherwise make ONE leaf whose near-field list holds every source leaf within (R_leaf + R_src) / theta — hundreds of
// thousands of particles for a single warp (measured: 10 ms tails on a 5 ms near-field launch once the padding particles of the
// ring field drift off the axis).  Dense regions are untouched (their leaves are 1-2 core sizes wide); cells whose particles
// carry no meaningful core size (probes: sigma = 1e-6, /root/reference/src/FLOWUnsteady_simulation.jl:572) are left alone.
constexpr double FMM_LEAF_SIGMAS = 4.0;
__device__ __forceinline__ bool fmm_sparse_leaf_splits(const FmmCell& cell, const double* __restrict__ rec) {
    if (rec == nullptr || cell.count < 1) return false;
    double m = 1.0e300;                      // min of 1/sigma^2 over the cell's (few) particles
    for (int s = 0; s < cell.count; ++s) m = fmin(m, rec[(size_t)(cell.start + s) * REC_REALS + 9]);
    const double smax2 = 1.0 / m;            // sigma_max^2
    const double R2 = cell.R * cell.R;
    return R2 > FMM_LEAF_SIGMAS * FMM_LEAF_SIGMAS * smax2 && smax2 * 16777216.0 > R2;   // sigma_max > R / 4096
}
__device__ __forceinline__ bool fmm_cell_splits(const FmmCell& cell, const uint64_t* __restrict__ keys, int ncrit,
                                                const int* __restrict__ hpre, int Lc, const double* __restrict__ rec) {
    if (cell.level >= FMM_MAXLEVEL) return false;
    if (hpre != nullptr && cell.level < Lc) {
        const uint64_t q = keys[cell.start] >> (3 * (FMM_MAXLEVEL - cell.level));
        const int sh = 3 * (Lc - cell.level);
        if (hpre[(q + 1) << sh] - hpre[q << sh] > ncrit) return true;
        return fmm_sparse_leaf_splits(cell, rec);   // a leaf of the global top tree: wholly owned, so the local particles are all
    }
    if (cell.count > ncrit) return true;
    return fmm_sparse_leaf_splits(cell, rec);
}

// For each cell of the level: number of non-empty children (0 if the cell is a leaf).
__global__ void fmm_split_count_kernel(const FmmCell* __restrict__ cells, int c0, int c1, const uint64_t* __restrict__ keys,
                                       int ncrit, int* __restrict__ nchild, const int* __restrict__ hpre = nullptr, int Lc = 0,
                                       const double* __restrict__ rec = nullptr) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const FmmCell cell = cells[c];
    int nc = 0;
    if (fmm_cell_splits(cell, keys, ncrit, hpre, Lc, rec)) {
        const int shift = 3 * (FMM_MAXLEVEL - cell.level - 1);
        const uint64_t prefix = (keys[cell.start] >> (shift + 3)) << 3;
        int lo = cell.start;
        for (int o = 0; o < 8; ++o) {
            int hi = o == 7 ? cell.start + cell.count
                            : lower_bound_key(keys, lo, cell.start + cell.count, (prefix | (uint64_t)(o + 1)) << shift);
            nc += hi > lo;
            lo = hi;
        }
    }
    nchild[c - c0] = nc;
}

__global__ void fmm_split_emit_kernel(FmmCell* __restrict__ cells, int c0, int c1, const uint64_t* __restrict__ keys,
                                      int ncrit, const int* __restrict__ child_off, int next0,
                                      const int* __restrict__ hpre = nullptr, int Lc = 0, const double* __restrict__ rec = nullptr) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    FmmCell cell = cells[c];
    if (!fmm_cell_splits(cell, keys, ncrit, hpre, Lc, rec)) {
        cells[c].child0 = -1;
        cells[c].nchild = 0;
        return;
    }
    const int shift = 3 * (FMM_MAXLEVEL - cell.level - 1);
    const uint64_t prefix = (keys[cell.start] >> (shift + 3)) << 3;
    int lo = cell.start, k = 0;
    const int base = next0 + child_off[c - c0];
    const double h = 0.5 * cell.R;
    for (int o = 0; o < 8; ++o) {
        int hi = o == 7 ? cell.start + cell.count
                        : lower_bound_key(keys, lo, cell.start + cell.count, (prefix | (uint64_t)(o + 1)) << shift);
        if (hi > lo) {
            FmmCell ch;
            ch.start = lo;
            ch.count = hi - lo;
            ch.parent = c;
            ch.child0 = -1;
            ch.nchild = 0;
            ch.level = cell.level + 1;
            ch.cx = cell.cx + ((o & 4) ? h : -h);   // key bit order: x, y, z (fmm_keys_kernel)
            ch.cy = cell.cy + ((o & 2) ? h : -h);
            ch.cz = cell.cz + ((o & 1) ? h : -h);
            ch.R = h;
            ch.smax = 0.0;
            ch.pad_ = 0.0;
            cells[base + k] = ch;
            ++k;
        }
        lo = hi;
    }
    cells[c].child0 = base;
    cells[c].nchild = k;
}

// ---- upward pass -----------------------------------------------------------------------------------------------------
// Multipoles: M[(cell * 3 + comp) * NM + a].  One warp per leaf, fixed-order shuffle reduction.
template <int P>
__global__ void fmm_p2m_kernel(const FmmCell* __restrict__ cells, int ncells, const double* __restrict__ rec,
                               double* __restrict__ M) {
    using Ops = FmmOps<P>;
    const int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (c >= ncells) return;
    const FmmCell cell = cells[c];
    if (cell.nchild != 0) return;
    const int lane = threadIdx.x & 31;
    double m0[Ops::NM], m1[Ops::NM], m2[Ops::NM];
#pragma unroll
    for (int a = 0; a < Ops::NM; ++a) m0[a] = m1[a] = m2[a] = 0.0;
    for (int s = lane; s < cell.count; s += 32) {
        const double* r = rec + (size_t)(cell.start + s) * REC_REALS;
        // record holds G' = -Gamma/(4 pi): expand psi = sum (Gamma/4pi)/r  ->  charges q = -G'
        Ops::p2m(r[0] - cell.cx, r[1] - cell.cy, r[2] - cell.cz, -r[3], -r[4], -r[5], m0, m1, m2);
    }
#pragma unroll
    for (int a = 0; a < Ops::NM; ++a) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            m0[a] += __shfl_xor_sync(0xffffffffu, m0[a], o);
            m1[a] += __shfl_xor_sync(0xffffffffu, m1[a], o);
            m2[a] += __shfl_xor_sync(0xffffffffu, m2[a], o);
        }
    }
    if (lane == 0) {
        double* out = M + (size_t)c * 3 * Ops::NM;
#pragma unroll
        for (int a = 0; a < Ops::NM; ++a) {
            out[a] = m0[a];
            out[Ops::NM + a] = m1[a];
            out[2 * Ops::NM + a] = m2[a];
        }
    }
}

// smax of the leaves (one thread per cell; record slot 9 holds 1/sigma^2) and of the inner cells of one level.
__global__ void fmm_smax_leaf_kernel(FmmCell* __restrict__ cells, int ncells, const double* __restrict__ rec) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= ncells) return;
    const FmmCell cell = cells[c];
    if (cell.nchild != 0) return;
    double m = 1.0e300;  // min of 1/sigma^2
    for (int s = 0; s < cell.count; ++s) m = fmin(m, rec[(size_t)(cell.start + s) * REC_REALS + 9]);
    cells[c].smax = rsqrt(m);
}
__global__ void fmm_smax_up_kernel(FmmCell* __restrict__ cells, int c0, int c1) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    const FmmCell cell = cells[c];
    if (cell.nchild == 0) return;
    double m = 0.0;
    for (int k = 0; k < cell.nchild; ++k) m = fmax(m, cells[cell.child0 + k].smax);
    cells[c].smax = m;
}

// One thread per (cell of the level, component): gather the children in order.
template <int P>
__global__ void fmm_m2m_kernel(const FmmCell* __restrict__ cells, int c0, int c1, double* __restrict__ M) {
    using Ops = FmmOps<P>;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = c0 + t / 3, comp = t % 3;
    if (c >= c1) return;
    const FmmCell cell = cells[c];
    if (cell.nchild == 0) return;
    double mp[Ops::NM];
#pragma unroll
    for (int a = 0; a < Ops::NM; ++a) mp[a] = 0.0;
    for (int k = 0; k < cell.nchild; ++k) {
        const FmmCell ch = cells[cell.child0 + k];
        double mc[Ops::NM];
        const double* src = M + ((size_t)(cell.child0 + k) * 3 + comp) * Ops::NM;
#pragma unroll
        for (int a = 0; a < Ops::NM; ++a) mc[a] = src[a];
        Ops::m2m(ch.cx - cell.cx, ch.cy - cell.cy, ch.cz - cell.cz, mc, mp);
    }
    double* dst = M + ((size_t)c * 3 + comp) * Ops::NM;
#pragma unroll
    for (int a = 0; a < Ops::NM; ++a) dst[a] = mp[a];
}

// ---- dual tree traversal (breadth-first over (target, source) cell pairs) ---------------------------------------------
struct FmmCounters {
    unsigned int next, m2l, p2p, overflow;
};

__device__ __forceinline__ void push_pair(uint64_t* __restrict__ list, unsigned int* __restrict__ counter, unsigned int cap,
                                          unsigned int* __restrict__ overflow, int a, int b) {
    unsigned int k = atomicAdd(counter, 1u);
    if (k < cap) list[k] = ((uint64_t)(unsigned int)a << 32) | (unsigned int)b; else atomicExch(overflow, 1u);
}

// Pairs are 64-bit keys (target << 32 | source) so the M2L / P2P lists sort by (target, source) without repacking.
__global__ void fmm_traverse_kernel(const FmmCell* __restrict__ cells, const uint64_t* __restrict__ frontier, unsigned int nfront,
                                    double theta, double nzs_factor, uint64_t* __restrict__ next, unsigned int cap_next, uint64_t* __restrict__ m2l,
                                    unsigned int cap_m2l, uint64_t* __restrict__ p2p, unsigned int cap_p2p,
                                    FmmCounters* __restrict__ cnt, const int* __restrict__ mine) {
    unsigned int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nfront) return;
    const int pi = (int)(frontier[t] >> 32), pj = (int)(frontier[t] & 0xffffffffu);
    const FmmCell ci = cells[pi], cj = cells[pj];
    const double dx = ci.cx - cj.cx, dy = ci.cy - cj.cy, dz = ci.cz - cj.cz;
    const double d2 = dx * dx + dy * dy + dz * dz;
    const double rs = ci.R + cj.R;
    // well separated (ExaFMM's multipole acceptance criterion); with nonzero_sigma the closest possible pair of points of the
    // two cells — the exact gap between the two cubes, not a bounding-sphere estimate, which is 3.7x too pessimistic for
    // neighbours-but-one — must also be nzs_factor core sizes of the SOURCE cell apart, so the singular far field is never used
    // where g(r/sigma) != 1
    bool well = rs * rs < theta * theta * d2;
    if (well && nzs_factor > 0.0) {
        const double gx = fmax(fabs(dx) - rs, 0.0), gy = fmax(fabs(dy) - rs, 0.0), gz = fmax(fabs(dz) - rs, 0.0);
        const double lim = nzs_factor * cj.smax;
        well = gx * gx + gy * gy + gz * gz > lim * lim;
    }
    if (well) {
        push_pair(m2l, &cnt->m2l, cap_m2l, &cnt->overflow, pi, pj);
    } else if (ci.nchild == 0 && cj.nchild == 0) {
        push_pair(p2p, &cnt->p2p, cap_p2p, &cnt->overflow, pi, cj.start);   // low word: first particle of the source leaf
    } else if (cj.nchild == 0 || (ci.nchild != 0 && ci.R >= cj.R)) {
        // multi-GPU: only target cells with one of this rank's leaves below them are followed (`mine`, nullptr = all)
        for (int k = 0; k < ci.nchild; ++k)
            if (!mine || mine[ci.child0 + k]) push_pair(next, &cnt->next, cap_next, &cnt->overflow, ci.child0 + k, pj);
    } else {
        for (int k = 0; k < cj.nchild; ++k) push_pair(next, &cnt->next, cap_next, &cnt->overflow, pi, cj.child0 + k);
    }
}

// off[c] = first index whose target is >= c (list sorted by (target, source)); off[ncells] = n
__global__ void fmm_list_offsets_kernel(const uint64_t* __restrict__ keys, unsigned int n, int ncells, unsigned int* __restrict__ off) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > ncells) return;
    unsigned int lo = 0, hi = n;
    const uint64_t v = (uint64_t)(unsigned int)c << 32;
    while (lo < hi) {
        unsigned int mid = (lo + hi) >> 1;
        if (keys[mid] < v) lo = mid + 1; else hi = mid;
    }
    off[c] = lo;
}

// ---- M2L: one warp per target cell; lane l takes sources l, l+32, ... of the cell's (sorted) list with private
//      accumulators in shared memory, then a fixed-order cross-lane sum.  L[(cell * 3 + comp) * NL + b].
template <int P>
__global__ void __launch_bounds__(32)
fmm_m2l_kernel(const FmmCell* __restrict__ cells, int ncells, const uint64_t* __restrict__ keys,
               const unsigned int* __restrict__ off, const double* __restrict__ M, double* __restrict__ L, FmmHalo halo) {
    using Ops = FmmOps<P>;
    extern __shared__ double sacc[];  // [3 * NL][32]  (coefficient-major: conflict-free per-lane columns)
    const int c = blockIdx.x;
    const unsigned int b0 = off[c], b1 = off[c + 1];
    if (b0 == b1) return;
    const int lane = threadIdx.x;
    for (int k = 0; k < 3 * Ops::NL; ++k) sacc[k * 32 + lane] = 0.0;
    const FmmCell ci = cells[c];
    for (unsigned int k = b0 + lane; k < b1; k += 32) {
        const int j = (int)(keys[k] & 0xffffffffu);
        const FmmCell cj = cells[j];
        double D[Ops::NL];
        Ops::dtensor(ci.cx - cj.cx, ci.cy - cj.cy, ci.cz - cj.cz, D);
#pragma unroll 1
        for (int comp = 0; comp < 3; ++comp) {
            double m[Ops::NM];
            const double* src = j >= halo.cell_split ? halo.M2 + ((size_t)halo.mslot[j] * 3 + comp) * Ops::NM
                                                     : M + ((size_t)j * 3 + comp) * Ops::NM;
#pragma unroll
            for (int a = 0; a < Ops::NM; ++a) m[a] = src[a];
            Ops::template m2l<32>(D, m, sacc + comp * Ops::NL * 32 + lane);   // this lane's private column
        }
    }
    __syncwarp();
    for (int k = lane; k < 3 * Ops::NL; k += 32) {
        double s = 0.0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) s += sacc[k * 32 + l];
        L[(size_t)c * 3 * Ops::NL + k] = s;
    }
}

// One thread per (cell of the level, component): add the parent's shifted expansion.
template <int P>
__global__ void fmm_l2l_kernel(const FmmCell* __restrict__ cells, int c0, int c1, double* __restrict__ L) {
    using Ops = FmmOps<P>;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int c = c0 + t / 3, comp = t % 3;
    if (c >= c1) return;
    const FmmCell cell = cells[c];
    if (cell.parent < 0) return;
    const FmmCell par = cells[cell.parent];
    const double* src = L + ((size_t)cell.parent * 3 + comp) * Ops::NL;
    double* dst = L + ((size_t)c * 3 + comp) * Ops::NL;
    Ops::l2l(cell.cx - par.cx, cell.cy - par.cy, cell.cz - par.cz, src, dst);
}

// ---- near field --------------------------------------------------------------------------------------------------
// P2P list entries are (target leaf cell << 32 | first particle of the source leaf), sorted; `runs` gives every entry's
// particle range.  Because the low word is a position in Morton order, consecutive entries are often adjacent ranges.
__global__ void fmm_leaf_counts_kernel(const FmmCell* __restrict__ cells, const int* __restrict__ leaves, int nleaves,
                                       int* __restrict__ count_at) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nleaves) {
        const FmmCell c = cells[leaves[k]];
        count_at[c.start] = c.count;
    }
}
__global__ void fmm_p2p_runs_kernel(const uint64_t* __restrict__ keys, unsigned int n, const int* __restrict__ count_at,
                                    int2* __restrict__ runs) {
    unsigned int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        const int start = (int)(keys[k] & 0xffffffffu);
        runs[k] = make_int2(start, count_at[start]);
    }
}

constexpr int LEAF_WARPS = 8;     // leaves per CTA (one warp each)
constexpr int LEAF_BATCH = 48;    // source records per staged batch (3.84 KB); two batches per warp (double buffer)
constexpr int LEAF_MIN_T = 4;     // fewest targets per pass -> at most 8 ways; LEAF_BATCH is a multiple of 2 * 8
static_assert(LEAF_BATCH % (2 * (32 / LEAF_MIN_T)) == 0, "a full batch must hold whole groups of 2 * ways records");
// Geometry of the UJ near-field kernel with the G table stored REP times (bank-conflict-free lookups, REP = 8): the 114 KB
// table leaves room for ONE CTA per SM, so it carries all 16 warps the 128 registers allow and stages 32-record batches.
template <int REP> struct LeafGeom { static constexpr int WARPS = LEAF_WARPS, BATCH = LEAF_BATCH; };
template <> struct LeafGeom<8> { static constexpr int WARPS = 16, BATCH = 32; };
static_assert(LeafGeom<8>::BATCH % (2 * (32 / LEAF_MIN_T)) == 0, "a full batch must hold whole groups of 2 * ways records");

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// The run list [b0, b1) of a leaf, read 32 entries at a time (lane l holds entry base + l, the following 32 are already in
// flight) and handed out by shuffle, so staging never waits on a dependent global load per run.
struct RunReader {
    const int2* runs;
    unsigned int b1, base;
    int2 cur, nxt;
    __device__ __forceinline__ int2 fetch(unsigned int k) const { return k < b1 ? runs[k] : make_int2(0, 0); }
    __device__ __forceinline__ void init(const int2* __restrict__ r, unsigned int b0, unsigned int b1_, int lane) {
        runs = r; b1 = b1_; base = b0;
        cur = fetch(b0 + lane);
        nxt = fetch(b0 + 32 + lane);
    }
    // k is warp-uniform, k < b1, and never decreases or skips between calls
    __device__ __forceinline__ int2 get(unsigned int k, int lane) {
        if (k - base >= 32u) {
            base += 32u;
            cur = nxt;
            nxt = fetch(base + 32u + lane);
        }
        const int j = (int)(k - base);
        return make_int2(__shfl_sync(0xffffffffu, cur.x, j), __shfl_sync(0xffffffffu, cur.y, j));
    }
};

// Issue the asynchronous copy (LDGSTS) of up to LEAF_BATCH source records of the run list [k, b1) into `slice`.
// (k, off) is the cursor: `off` records of run k are already consumed.  Returns the number of records in flight.
template <int BATCH = LEAF_BATCH>
__device__ __forceinline__ int stage_batch_async(RunReader& rr, unsigned int& k, unsigned int b1, int& off,
                                                 const double* __restrict__ rec, double* __restrict__ slice, int lane,
                                                 const FmmHalo& halo) {
    int n = 0;
    while (k < b1 && n < BATCH) {
        const int2 r = rr.get(k, lane);
        const int take = min(r.y - off, BATCH - n);
        const double* base = r.x >= halo.rec_split ? halo.rec2 + (size_t)(r.x - halo.rec_split + off) * REC_REALS
                                                   : rec + (size_t)(r.x + off) * REC_REALS;
        const double2* g2 = reinterpret_cast<const double2*>(base);
        double2* s2 = reinterpret_cast<double2*>(slice + (size_t)n * REC_REALS);
        for (int q = lane; q < take * (REC_REALS / 2); q += 32) cp_async16(s2 + q, g2 + q);
        n += take;
        off += take;
        if (off == r.y) { ++k; off = 0; }
    }
    cp_async_commit();
    return n;
}

// Round a landed batch of `ns` records up to a multiple of `mult` (= 2 * ways, a power of two <= 16) with copies of
// record 0 whose strength is zero (UJ record: G' in quads 1.y and 2; E_str record: quads 2..4): they contribute exactly 0
// on every branch and are never closer to a target than a real source, so the pair loop runs the same trip count on
// every lane (warp votes stay legal) and needs no mask.
template <bool ESTR>
__device__ __forceinline__ int pad_batch(double* __restrict__ buf, int ns, int mult, int lane) {
    const int np = (ns + mult - 1) & ~(mult - 1);
    if (np != ns) {
        double2* b2 = reinterpret_cast<double2*>(buf);
        for (int q = lane; q < (np - ns) * (REC_REALS / 2); q += 32) {
            const int part = q % (REC_REALS / 2);
            double2 v = b2[part];
            if (ESTR) {
                if (part >= 2) v = make_double2(0.0, 0.0);
            } else {
                if (part == 1) v.y = 0.0;
                if (part == 2) v = make_double2(0.0, 0.0);
            }
            b2[(ns + q / (REC_REALS / 2)) * (REC_REALS / 2) + part] = v;
        }
        __syncwarp();
    }
    return np;
}

// Lane mapping of one pass over a leaf's targets: T (a power of two) targets per pass and S = 32 / T "ways" that split
// the source loop (lane = way * T + t); the ways' partial sums are combined with xor-shuffles in a fixed order.  A leaf
// of `rem` < 32 remaining targets is taken as one pass with T = pow2ceil(rem), unless peeling off T/2 targets first leaves
// a remainder that fits a strictly smaller pass (18 targets: 16 x 2 ways, then 2 of 4 x 8 ways = 20/32 of the cost
// of one 32-target pass).  Each pass re-stages the leaf's sources (L2 hits); the split depends only on the count.
__device__ __forceinline__ int pow2ceil32(int v) {
    int t = 1;
    while (t < v) t <<= 1;
    return t;
}
__device__ __forceinline__ void leaf_pass(int rem, int& T, int& take) {
    if (rem >= 32) { T = 32; take = 32; return; }
    const int Tc = max(pow2ceil32(rem), LEAF_MIN_T);
    const int Tf = Tc >> 1;
    if (rem < Tc && Tf >= LEAF_MIN_T && max(pow2ceil32(rem - Tf), LEAF_MIN_T) < Tf) { T = Tf; take = Tf; }
    else { T = Tc; take = rem; }
}
__device__ __forceinline__ double xor_sum(double v, int T) {
    for (int o = 16; o >= T; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Two sources at a time for a whole warp (every lane must call: warp votes).  Inside an FMM near field nearly every pair
// is in the regularised range, so the first vote selects a straight-line block with two independent table evaluations.
template <int KERNEL, int REP>
__device__ __forceinline__ void uj_pair2_leaf(UJAcc& a, double tx, double ty, double tz, const double2* __restrict__ recA,
                                              const double2* __restrict__ recB, const double2* __restrict__ tab, int q) {
    const SrcCore sa = load_core(recA), sb = load_core(recB);
    double dxa = tx - sa.x, dya = ty - sa.y, dza = tz - sa.z;
    double dxb = tx - sb.x, dyb = ty - sb.y, dzb = tz - sb.z;
    double r2a = fma(dza, dza, fma(dya, dya, dxa * dxa));
    double r2b = fma(dzb, dzb, fma(dyb, dyb, dxb * dxb));
    if (KERNEL == K_SINGULAR) {
        const bool nza = nonzero_f64(r2a), nzb = nonzero_f64(r2b);
        double Aa, Ba, Ab, Bb;
        ab_singular(nza ? r2a : 1.0, Aa, Ba);
        ab_singular(nzb ? r2b : 1.0, Ab, Bb);
        uj_accumulate(a, dxa, dya, dza, sa.gx, sa.gy, sa.gz, nza ? Aa : 0.0, Ba);
        uj_accumulate(a, dxb, dyb, dzb, sb.gx, sb.gy, sb.gz, nzb ? Ab : 0.0, Bb);
    } else if (KERNEL == K_GAUSSIANERF) {
        const double2 qa3 = recA[3], qb3 = recB[3];
        const bool fa = __double2hiint(r2a) > __double2hiint(qa3.x), fb = __double2hiint(r2b) > __double2hiint(qb3.x);
        if (__all_sync(0xffffffffu, !fa && !fb)) {
            const double2 qa4 = recA[4], qb4 = recB[4];
            double Aa, Ba, Ab, Bb;
            ab_gauss_table_rep<REP>(tab, q, r2a * qa4.y, qa3.y, qa4.x, Aa, Ba);
            ab_gauss_table_rep<REP>(tab, q, r2b * qb4.y, qb3.y, qb4.x, Ab, Bb);
            Aa = nonzero_f64(r2a) ? Aa : 0.0;  // r == 0 skip (src/FLOWUnsteady_processing_force.jl:895)
            Ab = nonzero_f64(r2b) ? Ab : 0.0;
            uj_accumulate(a, dxa, dya, dza, sa.gx, sa.gy, sa.gz, Aa, Ba);
            uj_accumulate(a, dxb, dyb, dzb, sb.gx, sb.gy, sb.gz, Ab, Bb);
        } else if (__all_sync(0xffffffffu, fa && fb)) {
            double Aa, Ba, Ab, Bb;
            ab_singular(r2a, Aa, Ba);
            ab_singular(r2b, Ab, Bb);
            uj_accumulate(a, dxa, dya, dza, sa.gx, sa.gy, sa.gz, Aa, Ba);
            uj_accumulate(a, dxb, dyb, dzb, sb.gx, sb.gy, sb.gz, Ab, Bb);
        } else {
            uj_pair_general<KERNEL, REP>(a, dxa, dya, dza, r2a, sa, recA, tab, q);
            uj_pair_general<KERNEL, REP>(a, dxb, dyb, dzb, r2b, sb, recB, tab, q);
        }
    } else {
        uj_pair_general<KERNEL>(a, dxa, dya, dza, r2a, sa, recA, tab);
        uj_pair_general<KERNEL>(a, dxb, dyb, dzb, r2b, sb, recB, tab);
    }
}

// L2P + near-field P2P: one warp per leaf; warps are persistent and draw leaves from a global counter (leaf costs vary by
// two orders of magnitude, so a static 8-leaves-per-CTA split leaves a quarter of the warp slots idle).  Outputs in
// Morton order: sU[k * lds + i], sJ[k * lds + i].
template <int KERNEL, int P, int REP>
__global__ void __launch_bounds__(32 * LeafGeom<REP>::WARPS)
fmm_leaf_uj_kernel(const FmmCell* __restrict__ cells, const int* __restrict__ leaves, int nleaves, unsigned int* next_leaf,
                   const int2* __restrict__ runs, const unsigned int* __restrict__ p2p_off, const double* __restrict__ rec,
                   const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ sz,
                   const double* __restrict__ L, const double* __restrict__ gh_table, double* __restrict__ sU,
                   double* __restrict__ sJ, int64_t lds, FmmHalo halo) {
    using Ops = FmmOps<P>;
    constexpr int BATCH = LeafGeom<REP>::BATCH;
    // shared: [G table x REP (gaussianerf)] then per warp { [3 * NL padded even] local expansion, 2 x [BATCH * 10] records }
    extern __shared__ __align__(16) double smem[];
    constexpr int LPAD = (3 * Ops::NL + 1) & ~1;
    constexpr int TABD = KERNEL == K_GAUSSIANERF ? VPM_GG_DOUBLES * REP : 0;
    constexpr int WARPD = LPAD + 2 * BATCH * REC_REALS;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (KERNEL == K_GAUSSIANERF) {
        const double2* g2 = reinterpret_cast<const double2*>(gh_table);
        double2* t2 = reinterpret_cast<double2*>(smem);
        for (int k = threadIdx.x; k < TABD / 2; k += blockDim.x) t2[k] = g2[k / REP];
    }
    __syncthreads();
    const double2* tab = reinterpret_cast<const double2*>(smem);
    const int tq = lane & (REP - 1);   // this lane's copy of the table
    double* sL = smem + TABD + (size_t)warp * WARPD;
    double* buf0 = sL + LPAD;
    double* buf1 = buf0 + BATCH * REC_REALS;
  for (;;) {   // (body keeps its indentation: one leaf per trip)
    int leaf = 0;
    if (lane == 0) leaf = (int)atomicAdd(next_leaf, 1u);
    leaf = __shfl_sync(0xffffffffu, leaf, 0);
    if (leaf >= nleaves) break;    // whole warp leaves together (the only block-wide barrier is behind it)
    const int c = leaves[leaf];
    const FmmCell cell = cells[c];
    __syncwarp();                  // every lane is done with the previous leaf's expansion
    for (int k = lane; k < 3 * Ops::NL; k += 32) sL[k] = L[(size_t)c * 3 * Ops::NL + k];
    __syncwarp();
    const unsigned int b0 = p2p_off[c], b1 = p2p_off[c + 1];

    for (int t0 = 0; t0 < cell.count;) {
        int T, take;
        leaf_pass(cell.count - t0, T, take);
        const int S = 32 / T;
        const int t = lane & (T - 1), way = lane / T;
        const bool live = t < take;
        const int i = cell.start + t0 + (live ? t : 0);
        const double px = sx[i], py = sy[i], pz = sz[i];
        UJAcc a;
        acc_zero(a);
        double trace0 = 0.0;   // trace of the L2P part of J (the pair loop leaves j8 alone, UJAcc)
        if (way == 0) {   // far field: psi_n = local expansion n; U = curl psi, J = grad U
            double g[3][3], h[3][6];
#pragma unroll
            for (int n = 0; n < 3; ++n) Ops::l2p(px - cell.cx, py - cell.cy, pz - cell.cz, sL + n * Ops::NL, g[n], h[n]);
            a.u0 = g[2][1] - g[1][2];   // U_k = eps_kmn d_m psi_n
            a.u1 = g[0][2] - g[2][0];
            a.u2 = g[1][0] - g[0][1];
            // J[k + 3 l] = d_l U_k = eps_kmn d_l d_m psi_n ; h index: xx 0, xy 1, xz 2, yy 3, yz 4, zz 5
            a.j0 = h[2][1] - h[1][2]; a.j1 = h[0][2] - h[2][0]; a.j2 = h[1][0] - h[0][1];   // l = x: (xx, xy, xz) = 0, 1, 2
            a.j3 = h[2][3] - h[1][4]; a.j4 = h[0][4] - h[2][1]; a.j5 = h[1][1] - h[0][3];   // l = y: (yx, yy, yz) = 1, 3, 4
            a.j6 = h[2][4] - h[1][5]; a.j7 = h[0][5] - h[2][2]; a.j8 = h[1][2] - h[0][4];   // l = z: (zx, zy, zz) = 2, 4, 5
            trace0 = a.j0 + a.j4 + a.j8;
        }
        // near field: double-buffered batches of source records; way w takes records w, w + S, ... two at a time
        RunReader rr;
        rr.init(runs, b0, b1, lane);
        unsigned int k = b0;
        int off = 0, pb = 0;
        __syncwarp();
        int n_next = stage_batch_async<BATCH>(rr, k, b1, off, rec, buf0, lane, halo);
        while (true) {
            cp_async_wait_all();
            __syncwarp();
            int ns = n_next;
            if (ns == 0) break;
            n_next = stage_batch_async<BATCH>(rr, k, b1, off, rec, pb ? buf0 : buf1, lane, halo);
            ns = pad_batch<false>(pb ? buf1 : buf0, ns, 2 * S, lane);
            const double2* r2p = reinterpret_cast<const double2*>(pb ? buf1 : buf0) + way * (REC_REALS / 2);
            for (int s = 0; s < ns; s += 2 * S) {
                const double2* ra = r2p + s * (REC_REALS / 2);
                uj_pair2_leaf<KERNEL, REP>(a, px, py, pz, ra, ra + S * (REC_REALS / 2), tab, tq);
            }
            pb ^= 1;
        }
        if (S > 1) {
            a.u0 = xor_sum(a.u0, T); a.u1 = xor_sum(a.u1, T); a.u2 = xor_sum(a.u2, T);
            a.j0 = xor_sum(a.j0, T); a.j1 = xor_sum(a.j1, T); a.j2 = xor_sum(a.j2, T);
            a.j3 = xor_sum(a.j3, T); a.j4 = xor_sum(a.j4, T); a.j5 = xor_sum(a.j5, T);
            a.j6 = xor_sum(a.j6, T); a.j7 = xor_sum(a.j7, T);
            a.w0 = xor_sum(a.w0, T); a.w1 = xor_sum(a.w1, T); a.w2 = xor_sum(a.w2, T);
        }
        if (live && way == 0) {
            acc_close_trace(a, trace0);
            a.j1 -= a.w2; a.j2 += a.w1;
            a.j3 += a.w2; a.j5 -= a.w0;
            a.j6 -= a.w1; a.j7 += a.w0;
            double o[12] = {a.u0, a.u1, a.u2, a.j0, a.j1, a.j2, a.j3, a.j4, a.j5, a.j6, a.j7, a.j8};
#pragma unroll
            for (int q = 0; q < 3; ++q) sU[(size_t)q * lds + i] = o[q];
#pragma unroll
            for (int q = 0; q < 9; ++q) sJ[(size_t)q * lds + i] = o[3 + q];
        }
        t0 += take;
    }
  }
}

// Two E_str sources at a time (every lane must call: warp vote).  gaussianerf: both beyond T_FAR for all lanes -> skip.
template <int KERNEL>
__device__ __forceinline__ void estr_pair2_leaf(EAcc& a, double tx, double ty, double tz, const double2* __restrict__ recA,
                                                const double2* __restrict__ recB, const double* __restrict__ ztab) {
    if (KERNEL == K_GAUSSIANERF) {
        const double2 a0 = recA[0], a1 = recA[1], b0 = recB[0], b1 = recB[1];
        double dxa = tx - a0.x, dya = ty - a0.y, dza = tz - a1.x;
        double dxb = tx - b0.x, dyb = ty - b0.y, dzb = tz - b1.x;
        const double ta = fma(dza, dza, fma(dya, dya, dxa * dxa)) * a1.y;
        const double tb = fma(dzb, dzb, fma(dyb, dyb, dxb * dxb)) * b1.y;
        const bool fa = ta >= VPM_GT_TFAR, fb = tb >= VPM_GT_TFAR;
        if (__all_sync(0xffffffffu, fa && fb)) return;
        double za = zeta_gauss_table(ztab, fa ? 0.0 : ta);
        double zb = zeta_gauss_table(ztab, fb ? 0.0 : tb);
        za = fa ? 0.0 : za;
        zb = fb ? 0.0 : zb;
        const double2 a2 = recA[2], a3 = recA[3], a4 = recA[4];
        const double2 b2 = recB[2], b3 = recB[3], b4 = recB[4];
        a.a0 = fma(za, a2.x, a.a0); a.a1 = fma(za, a2.y, a.a1); a.a2 = fma(za, a3.x, a.a2);
        a.b0 = fma(za, a3.y, a.b0); a.b1 = fma(za, a4.x, a.b1); a.b2 = fma(za, a4.y, a.b2);
        a.a0 = fma(zb, b2.x, a.a0); a.a1 = fma(zb, b2.y, a.a1); a.a2 = fma(zb, b3.x, a.a2);
        a.b0 = fma(zb, b3.y, a.b0); a.b1 = fma(zb, b4.x, a.b1); a.b2 = fma(zb, b4.y, a.b2);
    } else {
        estr_pair<KERNEL>(a, tx, ty, tz, recA, ztab);
        estr_pair<KERNEL>(a, tx, ty, tz, recB, ztab);
    }
}

// Near-field E_str over the same leaf pairs (second pass; needs the converged J of targets and sources).
template <int KERNEL>
__global__ void __launch_bounds__(32 * LEAF_WARPS)
fmm_leaf_estr_kernel(const FmmCell* __restrict__ cells, const int* __restrict__ leaves, int nleaves, unsigned int* next_leaf,
                     const int2* __restrict__ runs, const unsigned int* __restrict__ p2p_off, const double* __restrict__ rec,
                     const double* __restrict__ sx, const double* __restrict__ sy, const double* __restrict__ sz,
                     const double* __restrict__ sJ, int64_t lds, int transposed, const double* __restrict__ z_table,
                     double* __restrict__ sE, FmmHalo halo) {
    extern __shared__ __align__(16) double smem[];   // [Z table (gaussianerf)] then per warp 2 x [LEAF_BATCH * 10] records
    constexpr int TABD = KERNEL == K_GAUSSIANERF ? ((VPM_GZ_NINT + 1) & ~1) : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (KERNEL == K_GAUSSIANERF)
        for (int k = threadIdx.x; k < VPM_GZ_NINT; k += blockDim.x) smem[k] = z_table[k];
    __syncthreads();
    const double* ztab = smem;
    double* buf0 = smem + TABD + (size_t)warp * (2 * LEAF_BATCH * REC_REALS);
    double* buf1 = buf0 + LEAF_BATCH * REC_REALS;
  for (;;) {   // persistent warps, one leaf per trip (see fmm_leaf_uj_kernel)
    int leaf = 0;
    if (lane == 0) leaf = (int)atomicAdd(next_leaf, 1u);
    leaf = __shfl_sync(0xffffffffu, leaf, 0);
    if (leaf >= nleaves) break;
    const int c = leaves[leaf];
    const FmmCell cell = cells[c];
    const unsigned int b0 = p2p_off[c], b1 = p2p_off[c + 1];
    for (int t0 = 0; t0 < cell.count;) {
        int T, take;
        leaf_pass(cell.count - t0, T, take);
        const int S = 32 / T;
        const int t = lane & (T - 1), way = lane / T;
        const bool live = t < take;
        const int i = cell.start + t0 + (live ? t : 0);
        const double px = sx[i], py = sy[i], pz = sz[i];
        EAcc a = {0, 0, 0, 0, 0, 0};
        RunReader rr;
        rr.init(runs, b0, b1, lane);
        unsigned int k = b0;
        int off = 0, pb = 0;
        __syncwarp();
        int n_next = stage_batch_async(rr, k, b1, off, rec, buf0, lane, halo);
        while (true) {
            cp_async_wait_all();
            __syncwarp();
            int ns = n_next;
            if (ns == 0) break;
            n_next = stage_batch_async(rr, k, b1, off, rec, pb ? buf0 : buf1, lane, halo);
            ns = pad_batch<true>(pb ? buf1 : buf0, ns, 2 * S, lane);
            const double2* r2p = reinterpret_cast<const double2*>(pb ? buf1 : buf0) + way * (REC_REALS / 2);
            for (int s = 0; s < ns; s += 2 * S) {
                const double2* ra = r2p + s * (REC_REALS / 2);
                estr_pair2_leaf<KERNEL>(a, px, py, pz, ra, ra + S * (REC_REALS / 2), ztab);
            }
            pb ^= 1;
        }
        if (S > 1) {
            a.a0 = xor_sum(a.a0, T); a.a1 = xor_sum(a.a1, T); a.a2 = xor_sum(a.a2, T);
            a.b0 = xor_sum(a.b0, T); a.b1 = xor_sum(a.b1, T); a.b2 = xor_sum(a.b2, T);
        }
        if (live && way == 0) {
            double Jp[9];
#pragma unroll
            for (int q = 0; q < 9; ++q) Jp[q] = sJ[(size_t)q * lds + i];
            double e0, e1, e2;
            if (transposed) {
                e0 = Jp[0] * a.a0 + Jp[1] * a.a1 + Jp[2] * a.a2;
                e1 = Jp[3] * a.a0 + Jp[4] * a.a1 + Jp[5] * a.a2;
                e2 = Jp[6] * a.a0 + Jp[7] * a.a1 + Jp[8] * a.a2;
            } else {
                e0 = Jp[0] * a.a0 + Jp[3] * a.a1 + Jp[6] * a.a2;
                e1 = Jp[1] * a.a0 + Jp[4] * a.a1 + Jp[7] * a.a2;
                e2 = Jp[2] * a.a0 + Jp[5] * a.a1 + Jp[8] * a.a2;
            }
            sE[0 * lds + i] = e0 - a.b0;
            sE[1 * lds + i] = e1 - a.b1;
            sE[2 * lds + i] = e2 - a.b2;
        }
        t0 += take;
    }
  }
}

// Leaves in Morton order: key = first particle of the leaf.
__global__ void fmm_leaf_starts_kernel(const FmmCell* __restrict__ cells, const int* __restrict__ leaves, int nleaves,
                                       int* __restrict__ starts) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < nleaves) starts[k] = cells[leaves[k]].start;
}

// First leaf (Morton order) whose first particle is >= target, for the two ends of a rank's share.
__global__ void fmm_leaf_range_kernel(const int* __restrict__ starts, int nleaves, int p_lo, int p_hi, int* __restrict__ out) {
    if (threadIdx.x >= 2) return;
    const int v = threadIdx.x == 0 ? p_lo : p_hi;
    int lo = 0, hi = nleaves;
    while (lo < hi) {
        int mid = (lo + hi) >> 1;
        if (starts[mid] < v) lo = mid + 1; else hi = mid;
    }
    out[threadIdx.x] = lo;
}

// mine[c] = 1 for every cell that is one of the given leaves or an ancestor of one
__global__ void fmm_flag_owned_kernel(const FmmCell* __restrict__ cells, const int* __restrict__ leaves, int nl, int* __restrict__ mine) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= nl) return;
    int c = leaves[k];
    while (c >= 0 && !mine[c]) {   // benign race: several leaves may mark the same ancestor
        mine[c] = 1;
        c = cells[c].parent;
    }
}

// leaves of the tree, in cell order
__global__ void fmm_mark_leaves_kernel(const FmmCell* __restrict__ cells, int ncells, int* __restrict__ flag) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells) flag[c] = cells[c].nchild == 0 ? 1 : 0;
}
__global__ void fmm_collect_leaves_kernel(const int* __restrict__ flag, const int* __restrict__ pos, int ncells,
                                          int* __restrict__ leaves) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c < ncells && flag[c]) leaves[pos[c]] = c;
}

// Morton order -> particle order: dst[k * ld + perm[i]] (+)= src[k * lds + i]
__global__ void fmm_scatter_kernel(const double* __restrict__ src, int64_t lds, int nrows, int64_t n, const int* __restrict__ perm,
                                   double* __restrict__ dst, int64_t ld, int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = perm[i];
    for (int k = 0; k < nrows; ++k) {
        double* d = dst + (size_t)k * ld + p;
        const double v = src[(size_t)k * lds + i];
        *d = accumulate ? *d + v : v;
    }
}

}  // namespace vpm
