// field_kernels.cuh — O(N) per-particle kernels of the hot path (all HBM-bound, coalesced SoA rows):
//   K6  AoS (43 x np column-major, the reference's matrix) <-> SoA import/export, source-record packing
//   K4  dynamic SFS coefficient (pseudo3level), clippings and controls          SURVEY.md A.5
//   K5  euler / rungekutta3 substep with the rVPM closure, core spreading        SURVEY.md A.6, A.8
//       and Pedrizzetti relaxation                                               SURVEY.md A.7
// Equations: /root/reference/docs/src/theory/rvpm.md:107-235 (governing equations), :264-296 (C_d), :363-367.
// Arithmetic is kept in the same operation order as oracle/vpm_oracle.c so the per-particle stages agree to
// round-off (FMA contraction is disabled for this translation unit's update kernels via explicit intrinsics).
#pragma once

#include "../../include/vpmb200.h"
#include "common.cuh"
#include "gauss_table.inc"

namespace vpm {

constexpr int PK_BT = 256;

// field-group table: first row and row count of each VPMB200_FM_* bit
__constant__ int c_group_first[13] = {F_X, F_GAMMA, F_SIGMA, F_VOL, F_CIRC, F_U, F_W, F_J, F_PSE, F_M, F_C, F_SFS, F_STATIC};
__constant__ int c_group_count[13] = {3, 3, 1, 1, 1, 3, 3, 9, 3, 9, 3, 3, 1};

__device__ __forceinline__ bool row_selected(int row, uint32_t mask) {
#pragma unroll
    for (int g = 0; g < 13; ++g)
        if ((mask >> g) & 1u)
            if (row >= c_group_first[g] && row < c_group_first[g] + c_group_count[g]) return true;
    return false;
}

// AoS -> SoA through a 32 x 33 shared tile so both sides are coalesced.  aos: particle-major, lda doubles per
// particle; particles [p0, p0 + n) of the AoS block land at SoA columns [dst0, dst0 + n).
__global__ void aos_to_soa_kernel(const double* __restrict__ aos, int64_t lda, int64_t n, double* __restrict__ soa,
                                  int64_t ld, int64_t dst0, uint32_t mask) {
    __shared__ double tile[32][33];
    const int64_t pbase = (int64_t)blockIdx.x * 32;
    const int fbase = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: particle within tile, threadIdx.x: field
        int64_t p = pbase + r;
        int f = fbase + threadIdx.x;
        tile[r][threadIdx.x] = (p < n && f < NFIELDS) ? aos[p * lda + f] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: field within tile, threadIdx.x: particle
        int f = fbase + r;
        int64_t p = pbase + threadIdx.x;
        if (p < n && f < NFIELDS && row_selected(f, mask)) soa[(size_t)f * ld + dst0 + p] = tile[threadIdx.x][r];
    }
}

__global__ void soa_to_aos_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, double* __restrict__ aos,
                                  int64_t lda, uint32_t mask) {
    __shared__ double tile[32][33];
    const int64_t pbase = (int64_t)blockIdx.x * 32;
    const int fbase = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: field, threadIdx.x: particle
        int f = fbase + r;
        int64_t p = pbase + threadIdx.x;
        tile[r][threadIdx.x] = (p < n && f < NFIELDS) ? soa[(size_t)f * ld + p] : 0.0;
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {  // r: particle, threadIdx.x: field
        int64_t p = pbase + r;
        int f = fbase + threadIdx.x;
        if (p < n && f < NFIELDS && row_selected(f, mask)) aos[p * lda + f] = tile[threadIdx.x][r];
    }
}

// rows [first, first + count) of particles [0, n) <- 0
__global__ void zero_rows_kernel(double* __restrict__ soa, int64_t ld, int64_t n, int first, int count) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int f = first; f < first + count; ++f) soa[(size_t)f * ld + i] = 0.0;
}

// swap-remove: column `last` -> column `i`
__global__ void move_column_kernel(double* __restrict__ soa, int64_t ld, int64_t dst, int64_t src) {
    int f = threadIdx.x;
    if (f < NFIELDS) soa[(size_t)f * ld + dst] = soa[(size_t)f * ld + src];
}

// Block-wide bounding box + max(T_FAR sigma^2) of a tile's real sources -> 10-double tile header (common.cuh).
// One CTA of TILE_SRC threads per tile.
__device__ __forceinline__ void write_tile_header(double* __restrict__ hdr, bool real, double x, double y, double z,
                                                  double rfar2, int nreal) {
    __shared__ double red[7][TILE_SRC / 32];
    const double big = 1.0e300;
    double v[7] = {real ? x : big, real ? -x : big, real ? y : big, real ? -y : big, real ? z : big, real ? -z : big,
                   real ? -rfar2 : big};
#pragma unroll
    for (int c = 0; c < 7; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] = fmin(v[c], __shfl_xor_sync(0xffffffffu, v[c], o));
    }
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 7; ++c) red[c][threadIdx.x >> 5] = v[c];
    __syncthreads();
    if (threadIdx.x < 10) {
        double out = 0.0;
        if (threadIdx.x < 7) {
            double m = red[threadIdx.x][0];
            for (int k = 1; k < TILE_SRC / 32; ++k) m = fmin(m, red[threadIdx.x][k]);
            out = (threadIdx.x & 1) || threadIdx.x == 6 ? -m : m;  // slots 1,3,5 are maxima, 6 is max rfar2
            if (threadIdx.x == 6 && nreal == 0) out = 0.0;
        } else if (threadIdx.x == 7) {
            out = (double)nreal;
        }
        hdr[threadIdx.x] = out;
    }
}

// UJ source tiles (common.cuh).  Grid: ntiles CTAs of TILE_SRC threads; slot i >= n repeats the tile's first real
// source position with zero strength.
__global__ void __launch_bounds__(TILE_SRC)
pack_uj_records_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, const int* __restrict__ perm,
                       double* __restrict__ rec) {
    const int64_t t0 = (int64_t)blockIdx.x * TILE_SRC;
    const int64_t islot = t0 + threadIdx.x;
    const bool real = islot < n;
    // slot -> particle: identity, or Morton order when the caller sorted (perm); t0 < n always
    const int64_t isrc = perm ? perm[real ? islot : t0] : (real ? islot : t0);
    const int64_t i = isrc;
    double* tile = rec + (size_t)blockIdx.x * TILE_DOUBLES;
    double2* r = reinterpret_cast<double2*>(tile + (size_t)threadIdx.x * REC_REALS);
    double x = soa[(size_t)(F_X + 0) * ld + isrc], y = soa[(size_t)(F_X + 1) * ld + isrc], z = soa[(size_t)(F_X + 2) * ld + isrc];
    double rfar2 = 0.0;
    if (real) {
        double gx = soa[(size_t)(F_GAMMA + 0) * ld + i], gy = soa[(size_t)(F_GAMMA + 1) * ld + i],
               gz = soa[(size_t)(F_GAMMA + 2) * ld + i];
        double sg = soa[(size_t)F_SIGMA * ld + i];
        double si = 1.0 / sg, si2 = si * si, si3 = si2 * si;
        rfar2 = VPM_GT_TFAR * (sg * sg);
        r[0] = make_double2(x, y);
        r[1] = make_double2(z, -CONST4 * gx);
        r[2] = make_double2(-CONST4 * gy, -CONST4 * gz);
        r[3] = make_double2(rfar2, si3);
        r[4] = make_double2(si3 * si2, si2);
    } else {
        r[0] = make_double2(x, y);
        r[1] = make_double2(z, 0.0);
        r[2] = make_double2(0.0, 0.0);
        r[3] = make_double2(VPM_GT_TFAR, 0.0);
        r[4] = make_double2(0.0, 1.0);
    }
    int64_t nreal = n - t0;
    write_tile_header(tile + TILE_HDR, real, x, y, z, rfar2, (int)(nreal > TILE_SRC ? TILE_SRC : nreal));
}

// E_str source tiles: c = zeta_norm / sigma^3, v = J^T Gamma (transposed) or J Gamma.  `cutoff` != 0 stores
// max(T_FAR sigma^2) in the header (kernels whose zeta vanishes beyond T_FAR); otherwise +inf-like (never skipped).
__global__ void __launch_bounds__(TILE_SRC)
pack_estr_records_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, const int* __restrict__ perm, int transposed,
                         double zeta_norm, int cutoff, double* __restrict__ rec) {
    const int64_t t0 = (int64_t)blockIdx.x * TILE_SRC;
    const int64_t islot = t0 + threadIdx.x;
    const bool real = islot < n;
    const int64_t isrc = perm ? perm[real ? islot : t0] : (real ? islot : t0);
    const int64_t i = isrc;
    double* tile = rec + (size_t)blockIdx.x * TILE_DOUBLES;
    double2* r = reinterpret_cast<double2*>(tile + (size_t)threadIdx.x * REC_REALS);
    double x = soa[(size_t)(F_X + 0) * ld + isrc], y = soa[(size_t)(F_X + 1) * ld + isrc], z = soa[(size_t)(F_X + 2) * ld + isrc];
    double rfar2 = 0.0;
    if (real) {
        double g0 = soa[(size_t)(F_GAMMA + 0) * ld + i], g1 = soa[(size_t)(F_GAMMA + 1) * ld + i],
               g2 = soa[(size_t)(F_GAMMA + 2) * ld + i];
        double sg = soa[(size_t)F_SIGMA * ld + i];
        double Jq[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) Jq[c] = soa[(size_t)(F_J + c) * ld + i];
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
        rfar2 = cutoff ? VPM_GT_TFAR * (sg * sg) : 1.0e300;
        r[0] = make_double2(x, y);
        r[1] = make_double2(z, si2);
        r[2] = make_double2(c * g0, c * g1);
        r[3] = make_double2(c * g2, c * v0);
        r[4] = make_double2(c * v1, c * v2);
    } else {
        r[0] = make_double2(x, y);
        r[1] = make_double2(z, 1.0);
        r[2] = make_double2(0.0, 0.0);
        r[3] = make_double2(0.0, 0.0);
        r[4] = make_double2(0.0, 0.0);
    }
    int64_t nreal = n - t0;
    write_tile_header(tile + TILE_HDR, real, x, y, z, rfar2, (int)(nreal > TILE_SRC ? TILE_SRC : nreal));
}

// zeta-pass source tiles: { x, y | z, 1/sigma^2 | c v0, c v1 | c v2, 0 | 0, 0 }, c = zeta_norm / sigma^3, v = 3 rows at `vec`
__global__ void __launch_bounds__(TILE_SRC)
pack_zeta_records_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, const double* __restrict__ vec, int64_t ldv,
                         double zeta_norm, int cutoff, double* __restrict__ rec) {
    const int64_t t0 = (int64_t)blockIdx.x * TILE_SRC;
    const int64_t i = t0 + threadIdx.x;
    const bool real = i < n;
    const int64_t isrc = real ? i : t0;
    double* tile = rec + (size_t)blockIdx.x * TILE_DOUBLES;
    double2* r = reinterpret_cast<double2*>(tile + (size_t)threadIdx.x * REC_REALS);
    double x = soa[(size_t)(F_X + 0) * ld + isrc], y = soa[(size_t)(F_X + 1) * ld + isrc], z = soa[(size_t)(F_X + 2) * ld + isrc];
    double rfar2 = 0.0;
    if (real) {
        double sg = soa[(size_t)F_SIGMA * ld + i];
        double si = 1.0 / sg, si2 = si * si;
        double c = zeta_norm * (si2 * si);
        rfar2 = cutoff ? VPM_GT_TFAR * (sg * sg) : 1.0e300;
        r[0] = make_double2(x, y);
        r[1] = make_double2(z, si2);
        r[2] = make_double2(c * vec[i], c * vec[ldv + i]);
        r[3] = make_double2(c * vec[2 * ldv + i], 0.0);
        r[4] = make_double2(0.0, 0.0);
    } else {
        r[0] = make_double2(x, y);
        r[1] = make_double2(z, 1.0);
        r[2] = make_double2(0.0, 0.0);
        r[3] = make_double2(0.0, 0.0);
        r[4] = make_double2(0.0, 0.0);
    }
    int64_t nreal = n - t0;
    write_tile_header(tile + TILE_HDR, real, x, y, z, rfar2, (int)(nreal > TILE_SRC ? TILE_SRC : nreal));
}

// ---- RBF conjugate gradient helpers (CoreSpreading spatial adaptation).  Vectors are 3 SoA rows; static
//      particles are fixed (their residual / direction entries are kept at zero). ---------------------------------------
// out[block * 4 + k] = sum over non-static i of a_k[i] * b_k[i]
__global__ void dot3_partials_kernel(const double* __restrict__ a, int64_t lda, const double* __restrict__ b, int64_t ldb,
                                     const double* __restrict__ stat, int64_t n, double* __restrict__ out) {
    __shared__ double red[3][8];
    double v[3] = {0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!(stat[i] > 0))
            for (int k = 0; k < 3; ++k) v[k] += a[(size_t)k * lda + i] * b[(size_t)k * ldb + i];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 3; ++c) red[c][threadIdx.x >> 5] = v[c];
    __syncthreads();
    if (threadIdx.x < 3) {
        double m = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); ++k) m += red[threadIdx.x][k];
        out[blockIdx.x * 4 + threadIdx.x] = m;
    }
}

// r = b - Ad (0 for statics), d = r
__global__ void cg_init_kernel(const double* __restrict__ b, const double* __restrict__ Ad, double* __restrict__ r,
                               double* __restrict__ d, int64_t ld, const double* __restrict__ stat, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool st = stat[i] > 0;
    for (int k = 0; k < 3; ++k) {
        double v = st ? 0.0 : b[(size_t)k * ld + i] - Ad[(size_t)k * ld + i];
        r[(size_t)k * ld + i] = v;
        d[(size_t)k * ld + i] = v;
    }
}

struct Vec3 {
    double v[3];
};

// x += alpha d ; r -= alpha Ad   (non-static)
__global__ void cg_update_kernel(double* __restrict__ x, double* __restrict__ r, const double* __restrict__ d,
                                 const double* __restrict__ Ad, int64_t ld, const double* __restrict__ stat, int64_t n, Vec3 alpha) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || stat[i] > 0) return;
    for (int k = 0; k < 3; ++k) {
        x[(size_t)k * ld + i] += alpha.v[k] * d[(size_t)k * ld + i];
        r[(size_t)k * ld + i] -= alpha.v[k] * Ad[(size_t)k * ld + i];
    }
}

// d = r + beta d   (non-static)
__global__ void cg_direction_kernel(double* __restrict__ d, const double* __restrict__ r, int64_t ld,
                                    const double* __restrict__ stat, int64_t n, Vec3 beta) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || stat[i] > 0) return;
    for (int k = 0; k < 3; ++k) d[(size_t)k * ld + i] = r[(size_t)k * ld + i] + beta.v[k] * d[(size_t)k * ld + i];
}

// max over non-static particles of sigma (block partials), and sigma <- sgm0 for them
__global__ void sigma_max_partials_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, double* __restrict__ out) {
    __shared__ double red[8];
    double m = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        if (!(soa[(size_t)F_STATIC * ld + i] > 0)) m = fmax(m, soa[(size_t)F_SIGMA * ld + i]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmax(m, red[k]);
        out[blockIdx.x] = m;
    }
}
__global__ void set_sigma_kernel(double* __restrict__ soa, int64_t ld, int64_t n, double sgm0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !(soa[(size_t)F_STATIC * ld + i] > 0)) soa[(size_t)F_SIGMA * ld + i] = sgm0;
}

// dst[k * ldd + i] = src[(row0 + k) * ld + perm[i]]  — SoA rows into Morton order
__global__ void gather_rows_kernel(const double* __restrict__ soa, int64_t ld, int row0, int nrows, int64_t n,
                                   const int* __restrict__ perm, double* __restrict__ dst, int64_t ldd) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int64_t p = perm[i];
    for (int k = 0; k < nrows; ++k) dst[(size_t)k * ldd + i] = soa[(size_t)(row0 + k) * ld + p];
}

// Sum the per-chunk partial rows of a source-split pair launch in chunk order.  partial row (c * ncomp + k) holds
// component k of chunk c; component k < nfirst goes to dstA[k * ldo + i], the rest to dstB[(k - nfirst) * ldo + i].
__global__ void reduce_partials_kernel(const double* __restrict__ partial, int nchunks, int ncomp, int64_t ldp, int64_t n,
                                       double* __restrict__ dstA, double* __restrict__ dstB, int nfirst, int64_t ldo,
                                       int accumulate) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    for (int k = 0; k < ncomp; ++k) {
        double* d = k < nfirst ? dstA + (size_t)k * ldo + i : dstB + (size_t)(k - nfirst) * ldo + i;
        double acc = accumulate ? *d : 0.0;
        for (int c = 0; c < nchunks; ++c) acc += partial[((size_t)c * ncomp + k) * ldp + i];
        *d = acc;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Per-particle stages.  `P(f)` addresses row f of this thread's particle.
// ------------------------------------------------------------------------------------------------------------
#define VPM_P(f) soa[(size_t)(f) * ld + i]

__device__ __forceinline__ void stretching(const double* __restrict__ soa, int64_t ld, int64_t i, int transposed,
                                           double G0, double G1, double G2, double S[3]) {
    double J[9];
#pragma unroll
    for (int c = 0; c < 9; ++c) J[c] = VPM_P(F_J + c);
    if (transposed) {
        S[0] = __dadd_rn(__dadd_rn(__dmul_rn(J[0], G0), __dmul_rn(J[1], G1)), __dmul_rn(J[2], G2));
        S[1] = __dadd_rn(__dadd_rn(__dmul_rn(J[3], G0), __dmul_rn(J[4], G1)), __dmul_rn(J[5], G2));
        S[2] = __dadd_rn(__dadd_rn(__dmul_rn(J[6], G0), __dmul_rn(J[7], G1)), __dmul_rn(J[8], G2));
    } else {
        S[0] = __dadd_rn(__dadd_rn(__dmul_rn(J[0], G0), __dmul_rn(J[3], G1)), __dmul_rn(J[6], G2));
        S[1] = __dadd_rn(__dadd_rn(__dmul_rn(J[1], G0), __dmul_rn(J[4], G1)), __dmul_rn(J[7], G2));
        S[2] = __dadd_rn(__dadd_rn(__dmul_rn(J[2], G0), __dmul_rn(J[5], G1)), __dmul_rn(J[8], G2));
    }
}

__device__ __forceinline__ double dot3(double a0, double a1, double a2, double b0, double b1, double b2) {
    return __dadd_rn(__dadd_rn(__dmul_rn(a0, b0), __dmul_rn(a1, b1)), __dmul_rn(a2, b2));
}

__device__ __forceinline__ double sgn(double x) { return (double)((x > 0) - (x < 0)); }

// sigma *= factor for the non-static particles (dynamic procedure test/domain filter switch)
__global__ void scale_sigma_kernel(double* __restrict__ soa, int64_t ld, int64_t n, double factor, int divide) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    VPM_P(F_SIGMA) = divide ? VPM_P(F_SIGMA) / factor : VPM_P(F_SIGMA) * factor;
}

// test-filter results: M[:,1] = S, M[:,2] = SFS, rest of M = 0
__global__ void dyn_store_test_kernel(double* __restrict__ soa, int64_t ld, int64_t n, int transposed) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    double S[3];
    stretching(soa, ld, i, transposed, VPM_P(F_GAMMA), VPM_P(F_GAMMA + 1), VPM_P(F_GAMMA + 2), S);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        VPM_P(F_M + k) = S[k];
        VPM_P(F_M + 3 + k) = VPM_P(F_SFS + k);
        VPM_P(F_M + 6 + k) = 0.0;
    }
}

// domain-filter results, C_d with Lagrangian averaging and clamps, flush M   
__global__ void dyn_coeff_kernel(double* __restrict__ soa, int64_t ld, int64_t n, int transposed, double alpha,
                                 double rlxf, double minC, double maxC, int force_positive, double zeta0) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    double G0 = VPM_P(F_GAMMA), G1 = VPM_P(F_GAMMA + 1), G2 = VPM_P(F_GAMMA + 2);
    double S[3];
    stretching(soa, ld, i, transposed, G0, G1, G2, S);
    double M1[3], M2[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        M1[k] = __dsub_rn(VPM_P(F_M + k), S[k]);
        M2[k] = __dsub_rn(VPM_P(F_M + 3 + k), VPM_P(F_SFS + k));
    }
    double sg = VPM_P(F_SIGMA);
    double nume = dot3(M1[0], M1[1], M1[2], G0, G1, G2);
    nume = __dmul_rn(nume, __dsub_rn(__dmul_rn(3.0, alpha), 2.0));
    double deno = dot3(M2[0], M2[1], M2[2], G0, G1, G2);
    deno = deno / (zeta0 / __dmul_rn(__dmul_rn(sg, sg), sg));
    double C1 = VPM_P(F_C + 1), C2 = VPM_P(F_C + 2);
    if (C2 == 0) {
        C2 = deno;
        if (C2 == 0) C2 = 2.220446049250313e-16;
    }
    nume = __dadd_rn(__dmul_rn(rlxf, nume), __dmul_rn(__dsub_rn(1.0, rlxf), C1));
    deno = __dadd_rn(__dmul_rn(rlxf, deno), __dmul_rn(__dsub_rn(1.0, rlxf), C2));
    if (fabs(nume / deno) > maxC) {
        if (fabs(deno) < fabs(C2)) deno = sgn(deno) * fabs(C2);
        nume = sgn(nume) * fabs(deno) * maxC;
    } else if (fabs(nume / deno) < minC) {
        nume = sgn(nume) * fabs(deno) * minC;
    }
    double C0 = nume / deno;
    if (force_positive) C0 = fabs(C0);
    VPM_P(F_C) = C0;
    VPM_P(F_C + 1) = nume;
    VPM_P(F_C + 2) = deno;
#pragma unroll
    for (int k = 0; k < 9; ++k) VPM_P(F_M + k) = 0.0;
}

__global__ void const_coeff_kernel(double* __restrict__ soa, int64_t ld, int64_t n, double Cs) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    VPM_P(F_C) = Cs;
}

// clipping_backscatter, control_directional, control_magnitude — applied in that order per particle; each only
// touches its own particle, so the reference's three sweeps fuse into one.
__global__ void clip_control_kernel(double* __restrict__ soa, int64_t ld, int64_t n, int clippings, int controls,
                                    double f, double zeta0, double t, int64_t nt) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    double G0 = VPM_P(F_GAMMA), G1 = VPM_P(F_GAMMA + 1), G2 = VPM_P(F_GAMMA + 2);
    double E0 = VPM_P(F_SFS), E1 = VPM_P(F_SFS + 1), E2 = VPM_P(F_SFS + 2);
    double C = VPM_P(F_C);
    if (clippings & VPMB200_CLIP_BACKSCATTER) {
        if (__dmul_rn(C, dot3(G0, G1, G2, E0, E1, E2)) < 0) {
            C *= 0;
            VPM_P(F_C) = C;
        }
    }
    if (controls & VPMB200_CTRL_DIRECTIONAL) {
        double aux = dot3(E0, E1, E2, G0, G1, G2);
        aux = aux / dot3(G0, G1, G2, G0, G1, G2);
        E0 = __dadd_rn(E0, __dadd_rn(-E0, __dmul_rn(aux, G0)));
        E1 = __dadd_rn(E1, __dadd_rn(-E1, __dmul_rn(aux, G1)));
        E2 = __dadd_rn(E2, __dadd_rn(-E2, __dmul_rn(aux, G2)));
    }
    if ((controls & VPMB200_CTRL_MAGNITUDE) && nt != 0 && C != 0) {
        double deltat = t / (double)nt;
        double sg = VPM_P(F_SIGMA);
        double aux = dot3(E0, E1, E2, G0, G1, G2);
        aux = aux / dot3(G0, G1, G2, G0, G1, G2);
        aux = __dsub_rn(aux, __dmul_rn(__dadd_rn(1.0, __dmul_rn(3.0, f)), (zeta0 / __dmul_rn(__dmul_rn(sg, sg), sg))) / deltat / C);
        if (aux > 0) {
            E0 = __dadd_rn(E0, __dmul_rn(-aux, G0));
            E1 = __dadd_rn(E1, __dmul_rn(-aux, G1));
            E2 = __dadd_rn(E2, __dmul_rn(-aux, G2));
        }
    }
    if (controls) {
        VPM_P(F_SFS) = E0;
        VPM_P(F_SFS + 1) = E1;
        VPM_P(F_SFS + 2) = E2;
    }
}

struct UpdateParams {
    double a, b, dt;
    double Uinf0, Uinf1, Uinf2;
    double f, g, zeta0, nu, rlxf;
    int transposed, viscous, relaxation;
    int euler;        // 1: Euler step — no q-storage, relaxation (if relax_inline) before core spreading
    int relax_inline;
};

__device__ __forceinline__ void relax_gamma(const double* __restrict__ soa, int64_t ld, int64_t i, int relaxation,
                                            double rlxf, double& G0, double& G1, double& G2) {
    // omega = curl u from J; J[i,j] at i + 3 j
    double w1 = __dsub_rn(VPM_P(F_J + 2 + 3 * 1), VPM_P(F_J + 1 + 3 * 2));
    double w2 = __dsub_rn(VPM_P(F_J + 0 + 3 * 2), VPM_P(F_J + 2 + 3 * 0));
    double w3 = __dsub_rn(VPM_P(F_J + 1 + 3 * 0), VPM_P(F_J + 0 + 3 * 1));
    double nrmw = sqrt(dot3(w1, w2, w3, w1, w2, w3));
    double nrmG = sqrt(dot3(G0, G1, G2, G0, G1, G2));
    double omr = __dsub_rn(1.0, rlxf);
    double b2 = 1.0;
    if (relaxation == VPMB200_RELAX_CORRECTEDPEDRIZZETTI)
        b2 = __dsub_rn(1.0, __dmul_rn(__dmul_rn(__dmul_rn(2.0, omr), rlxf),
                                      __dsub_rn(1.0, dot3(G0, G1, G2, w1, w2, w3) / __dmul_rn(nrmG, nrmw))));
    G0 = __dadd_rn(__dmul_rn(omr, G0), __dmul_rn(__dmul_rn(rlxf, nrmG), w1) / nrmw);
    G1 = __dadd_rn(__dmul_rn(omr, G1), __dmul_rn(__dmul_rn(rlxf, nrmG), w2) / nrmw);
    G2 = __dadd_rn(__dmul_rn(omr, G2), __dmul_rn(__dmul_rn(rlxf, nrmG), w3) / nrmw);
    if (relaxation == VPMB200_RELAX_CORRECTEDPEDRIZZETTI) {
        double sb = sqrt(b2);
        G0 /= sb;
        G1 /= sb;
        G2 /= sb;
    }
}

// One low-storage substep + core spreading (A.8) for every non-static particle.
__global__ void update_kernel(double* __restrict__ soa, int64_t ld, int64_t n, UpdateParams p) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    const double a = p.a, b = p.b, dt = p.dt;
    const double Uinf[3] = {p.Uinf0, p.Uinf1, p.Uinf2};
    double G[3] = {VPM_P(F_GAMMA), VPM_P(F_GAMMA + 1), VPM_P(F_GAMMA + 2)};
    double E[3] = {VPM_P(F_SFS), VPM_P(F_SFS + 1), VPM_P(F_SFS + 2)};
    const double C = VPM_P(F_C);
    // position
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double q = p.euler ? 0.0 : VPM_P(F_M + k);
        q = __dadd_rn(__dmul_rn(a, q), __dmul_rn(dt, __dadd_rn(VPM_P(F_U + k), Uinf[k])));
        if (!p.euler) VPM_P(F_M + k) = q;
        VPM_P(F_X + k) = __dadd_rn(VPM_P(F_X + k), __dmul_rn(b, q));
    }
    double S[3];
    stretching(soa, ld, i, p.transposed, G[0], G[1], G[2], S);
    double sg = VPM_P(F_SIGMA);
    double sg3z = __dmul_rn(__dmul_rn(sg, sg), sg) / p.zeta0;
    double Z = __dmul_rn(__dadd_rn(p.f, p.g) / __dadd_rn(1.0, __dmul_rn(3.0, p.f)), dot3(S[0], S[1], S[2], G[0], G[1], G[2]));
    double CE = __dadd_rn(__dadd_rn(__dmul_rn(__dmul_rn(C, E[0]), G[0]), __dmul_rn(__dmul_rn(C, E[1]), G[1])),
                          __dmul_rn(__dmul_rn(C, E[2]), G[2]));
    Z = __dsub_rn(Z, __dmul_rn(__dmul_rn(p.f / __dadd_rn(1.0, __dmul_rn(3.0, p.f)), CE), sg3z));
    Z = Z / dot3(G[0], G[1], G[2], G[0], G[1], G[2]);
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        double q = p.euler ? 0.0 : VPM_P(F_M + 3 + k);
        double rhs = __dsub_rn(__dsub_rn(S[k], __dmul_rn(__dmul_rn(3.0, Z), G[k])), __dmul_rn(__dmul_rn(C, E[k]), sg3z));
        q = __dadd_rn(__dmul_rn(a, q), __dmul_rn(dt, rhs));
        if (!p.euler) VPM_P(F_M + 3 + k) = q;
        G[k] = __dadd_rn(G[k], __dmul_rn(b, q));
    }
    {
        double q = p.euler ? 0.0 : VPM_P(F_M + 7);
        q = __dsub_rn(__dmul_rn(a, q), __dmul_rn(dt, __dmul_rn(sg, Z)));
        if (!p.euler) VPM_P(F_M + 7) = q;
        sg = __dadd_rn(sg, __dmul_rn(b, q));
    }
    if (p.euler && p.relax_inline && p.relaxation != VPMB200_RELAX_NONE)
        relax_gamma(soa, ld, i, p.relaxation, p.rlxf, G[0], G[1], G[2]);
    if (p.viscous == VPMB200_VISCOUS_CORESPREADING) {
        if (p.euler) {
            sg = sqrt(__dadd_rn(__dmul_rn(sg, sg), __dmul_rn(__dmul_rn(2.0, p.nu), dt)));
        } else {
            double q = __dadd_rn(__dmul_rn(a, VPM_P(F_M + 6)), __dmul_rn(__dmul_rn(dt, 2.0), p.nu));
            VPM_P(F_M + 6) = q;
            sg = sqrt(__dadd_rn(__dmul_rn(sg, sg), __dmul_rn(b, q)));
        }
    }
    VPM_P(F_GAMMA) = G[0];
    VPM_P(F_GAMMA + 1) = G[1];
    VPM_P(F_GAMMA + 2) = G[2];
    VPM_P(F_SIGMA) = sg;
}

__global__ void relax_kernel(double* __restrict__ soa, int64_t ld, int64_t n, int relaxation, double rlxf) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
    double G0 = VPM_P(F_GAMMA), G1 = VPM_P(F_GAMMA + 1), G2 = VPM_P(F_GAMMA + 2);
    relax_gamma(soa, ld, i, relaxation, rlxf, G0, G1, G2);
    VPM_P(F_GAMMA) = G0;
    VPM_P(F_GAMMA + 1) = G1;
    VPM_P(F_GAMMA + 2) = G2;
}

// rungekutta3 resets its q-storage M for the non-static particles before the first substep
__global__ void zero_m_kernel(double* __restrict__ soa, int64_t ld, int64_t n) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (VPM_P(F_STATIC) > 0) return;
#pragma unroll
    for (int k = 0; k < 9; ++k) VPM_P(F_M + k) = 0.0;
}

// number of non-finite entries among X, Gamma, sigma
__global__ void count_nonfinite_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, unsigned long long* out) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned int bad = 0;
    if (i < n)
        for (int f = 0; f < 7; ++f) bad += !isfinite(VPM_P(f));
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0 && bad) atomicAdd(out, (unsigned long long)bad);
}

// ---- wake treatments (FLOWUnsteady's remove_particles_* runtime functions, src/FLOWUnsteady_processing.jl:50-187) as a
//      stream compaction that reproduces the ORDER the reference's sequential loop leaves behind.  That loop walks
//      i = np..1 and vpm.remove_particle(i) moves the current last particle into slot i.  Tracking one survivor through
//      it: its index inside the list of survivors above the cursor grows by one per step and wraps to 0 exactly when it
//      is the last element and the cursor sits on a removed slot — it then lands in that hole.  With R(p) = removed slots
//      below p and C(p) = survivors at or above p, a particle (virtually) inserted at p next wraps at the removed slot of
//      ascending rank R(p) - C(p), or never if that is negative.  The final slot is the fixed point of that map, found by
//      pointer doubling (tests replay the reference loop line by line).
struct RemoveCriterion {
    int kind;          // VPMB200_REMOVE_*
    double p[9];
};

__global__ void keep_flags_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, RemoveCriterion c, int* __restrict__ keep) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    bool k = true;
    if (c.kind == VPMB200_REMOVE_STRENGTH) {          // keep iff minGamma2 <= |Gamma|^2 <= maxGamma2   (:52-56)
        double g0 = VPM_P(F_GAMMA), g1 = VPM_P(F_GAMMA + 1), g2 = VPM_P(F_GAMMA + 2);
        double m = g0 * g0 + g1 * g1 + g2 * g2;
        k = (c.p[0] <= m) && (m <= c.p[1]);
    } else if (c.kind == VPMB200_REMOVE_SIGMA) {      // keep iff minsigma <= sigma <= maxsigma        (:98-100)
        double sg = VPM_P(F_SIGMA);
        k = (c.p[0] <= sg) && (sg <= c.p[1]);
    } else if (c.kind == VPMB200_REMOVE_BOX) {        // remove if X - O is outside [Pmin, Pmax]       (:130-138)
        double x = VPM_P(F_X) - c.p[6], y = VPM_P(F_X + 1) - c.p[7], z = VPM_P(F_X + 2) - c.p[8];
        k = !((x < c.p[0] || x > c.p[3]) || (y < c.p[1] || y > c.p[4]) || (z < c.p[2] || z > c.p[5]));
    } else if (c.kind == VPMB200_REMOVE_SPHERE) {     // remove if |X - centre|^2 > Rsphere2            (:171-178)
        double x = VPM_P(F_X) - c.p[1], y = VPM_P(F_X + 1) - c.p[2], z = VPM_P(F_X + 2) - c.p[3];
        k = !(x * x + y * y + z * z > c.p[0]);
    }
    keep[i] = k ? 1 : 0;
}

// removed slots listed by ascending rank
__global__ void list_removed_kernel(const int* __restrict__ keep, const int* __restrict__ kept_before, int64_t n,
                                    int* __restrict__ removedpos) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && !keep[i]) removedpos[i - kept_before[i]] = (int)i;
}

// f(p): the hole a particle inserted at p falls into next (p itself when it never wraps again)
__global__ void wrap_map_kernel(const int* __restrict__ kept_before, int64_t n, int64_t K, const int* __restrict__ removedpos,
                                int* __restrict__ f) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int64_t R = p - kept_before[p], C = K - kept_before[p];
    f[p] = R - C >= 0 ? removedpos[R - C] : (int)p;
}

__global__ void wrap_double_kernel(const int* __restrict__ f, int64_t n, int* __restrict__ g, int* __restrict__ changed) {
    int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const int a = f[p], b = f[a];
    g[p] = b;
    if (a != b) *changed = 1;
}

// survivors move to their final slot (destinations are holes or the particle's own slot: no read/write overlap)
__global__ void move_survivors_kernel(double* __restrict__ soa, int64_t ld, const int* __restrict__ keep, int64_t n,
                                      const int* __restrict__ dest) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n || !keep[i]) return;
    const int64_t d = dest[i];
    if (d == i) return;
    for (int f = 0; f < NFIELDS; ++f) soa[(size_t)f * ld + d] = soa[(size_t)f * ld + i];
}

// ---- monitors: per-block partial sums (fixed order) of the quantities vpm.monitor_enstrophy / vpm.monitor_Cd report
//      (src/FLOWUnsteady_monitors.jl:614,697).  out[block * 8 + k]: 0 enstrophy 0.5 sum Gamma.omega, 1 sum C_d over
//      particles with C_d != 0, 2 sum C_d^2, 3 count C_d != 0, 4 number of static particles, 5 sum |Gamma|, 6, 7 unused.
__global__ void monitor_partials_kernel(const double* __restrict__ soa, int64_t ld, int64_t n, double* __restrict__ out) {
    __shared__ double red[8][8];
    double v[6] = {0, 0, 0, 0, 0, 0};
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        double g0 = VPM_P(F_GAMMA), g1 = VPM_P(F_GAMMA + 1), g2 = VPM_P(F_GAMMA + 2);
        double w0 = VPM_P(F_J + 5) - VPM_P(F_J + 7), w1 = VPM_P(F_J + 6) - VPM_P(F_J + 2), w2 = VPM_P(F_J + 1) - VPM_P(F_J + 3);
        v[0] += 0.5 * (g0 * w0 + g1 * w1 + g2 * w2);
        double C = VPM_P(F_C);
        if (C != 0) { v[1] += C; v[2] += C * C; v[3] += 1.0; }
        if (VPM_P(F_STATIC) > 0) v[4] += 1.0;
        v[5] += sqrt(g0 * g0 + g1 * g1 + g2 * g2);
    }
#pragma unroll
    for (int c = 0; c < 6; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] += __shfl_xor_sync(0xffffffffu, v[c], o);
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 6; ++c) red[c][threadIdx.x >> 5] = v[c];
    __syncthreads();
    if (threadIdx.x < 8) {
        double m = 0.0;
        if (threadIdx.x < 6)
            for (int k = 0; k < (int)(blockDim.x >> 5); ++k) m += red[threadIdx.x][k];
        out[blockIdx.x * 8 + threadIdx.x] = m;
    }
}

#undef VPM_P

}  // namespace vpm
