// uj_direct.cuh — K1: tiled O(N^2) regularised Biot-Savart P2P, velocity U and its gradient J.
//
// Replaces FLOWVPM's UJ_direct (selected through `vpm_UJ=vpm.UJ_direct`,
// /root/reference/src/FLOWUnsteady_simulation.jl:38; invoked as `pfield.UJ(pfield)` at :544 and at
// src/FLOWUnsteady_processing_force.jl:238).  Pair arithmetic follows SURVEY.md A.2, whose U part is
// pinned by the in-tree P2P src/FLOWUnsteady_processing_force.jl:889-905 (dx = x_target - x_source, sigma of
// the SOURCE, r != 0 skip, 1/(4 pi) literal).
//
// B200 mapping (FP64 vector pipe is the roofline; no tensor cores — this is not a contraction):
//   * one target per thread (U(3) + J(9) + delta-term(3) accumulators in registers), BT targets per CTA;
//   * sources stream through shared memory in TILE_SRC-record tiles, double-buffered with 1-D bulk TMA
//     copies (cp.async.bulk -> UBLKCP) completing on mbarriers; every lane reads the same record
//     (LDS.128 broadcast), so HBM/L2 traffic is 80 B per BT interactions;
//   * the pair math is re-derived in t = r^2/sigma^2 so the regularised (near) branch needs no sqrt, erf,
//     exp or division: g/r^3 = G(t)/sigma^3 and (g'/(sigma r) - 3 g/r^2)/r^3 = H(t)/sigma^5 with G, H
//     entire functions of t evaluated from a shared-memory piecewise-polynomial table
//     (tools/gen_tables.py, max rel. error 2e-15; H = 2 dG/dt so one polynomial gives both).  For t >= T_FAR the Gaussian-erf kernel equals the
//     singular kernel to < 2^-56, and a warp-uniform vote takes the branch with one MUFU.RSQ64H + 5 DFMA;
//   * -1/(4 pi) is folded into the source's Gamma, the antisymmetric (Kronecker-delta) part of J is
//     accumulated as a 3-vector and expanded once at the end, and per-tile partial sums are added to the
//     running totals (pairwise-style summation keeps the 1e6-term sums inside the 1e-12 parity budget).
#pragma once

#include "common.cuh"
#include "gauss_table.inc"

namespace vpm {

constexpr double MAGIC_RINT = 6755399441055744.0;  // 1.5 * 2^52: fma(t, 1/W, MAGIC) puts rint(t/W) in the low word

// J of a pair is B (c x dx^T) + (antisymmetric delta term) with c = dx x G' perpendicular to dx, so its trace vanishes
// identically (div u = 0): the pair loops do not accumulate j8; it is recovered once at the end as -(j0 + j4)
// (acc_close_trace).  That is one FP64 instruction less per interaction.
struct UJAcc {
    double u0, u1, u2;
    double j0, j1, j2, j3, j4, j5, j6, j7, j8;  // J[i + 3 j]; j8 is NOT touched by uj_accumulate
    double w0, w1, w2;                           // sum A * G'  (delta term, expanded at the end)
};

__device__ __forceinline__ void acc_zero(UJAcc& a) {
    a.u0 = a.u1 = a.u2 = 0.0;
    a.j0 = a.j1 = a.j2 = a.j3 = a.j4 = a.j5 = a.j6 = a.j7 = a.j8 = 0.0;
    a.w0 = a.w1 = a.w2 = 0.0;
}
__device__ __forceinline__ void acc_add(UJAcc& t, const UJAcc& a) {
    t.u0 += a.u0; t.u1 += a.u1; t.u2 += a.u2;
    t.j0 += a.j0; t.j1 += a.j1; t.j2 += a.j2; t.j3 += a.j3; t.j4 += a.j4;
    t.j5 += a.j5; t.j6 += a.j6; t.j7 += a.j7;
    t.w0 += a.w0; t.w1 += a.w1; t.w2 += a.w2;
}

// Common tail of one interaction once A = g/r^3 and B = (g'/(sigma r) - 3 g/r^2)/r^3 are known.
__device__ __forceinline__ void uj_accumulate(UJAcc& a, double dx, double dy, double dz, double gx, double gy,
                                              double gz, double A, double B) {
    // c = dx x G'   (G' = -Gamma/4pi, so A*c = g K x Gamma)
    double c0 = fma(dy, gz, -dz * gy);
    double c1 = fma(dz, gx, -dx * gz);
    double c2 = fma(dx, gy, -dy * gx);
    a.u0 = fma(A, c0, a.u0);
    a.u1 = fma(A, c1, a.u1);
    a.u2 = fma(A, c2, a.u2);
    double b0 = B * c0, b1 = B * c1, b2 = B * c2;
    a.j0 = fma(b0, dx, a.j0); a.j1 = fma(b1, dx, a.j1); a.j2 = fma(b2, dx, a.j2);
    a.j3 = fma(b0, dy, a.j3); a.j4 = fma(b1, dy, a.j4); a.j5 = fma(b2, dy, a.j5);
    a.j6 = fma(b0, dz, a.j6); a.j7 = fma(b1, dz, a.j7);   // j8: see UJAcc
    a.w0 = fma(A, gx, a.w0);
    a.w1 = fma(A, gy, a.w1);
    a.w2 = fma(A, gz, a.w2);
}

// Far-field flavour used by K1's all-far tiles: B3 = B / (-3) = 1/r^5 (one multiply less than B); the tile's J partial
// sums are scaled by -3 when they are added to the running totals (acc_add_far), which costs nothing extra.
__device__ __forceinline__ void uj_accumulate_far(UJAcc& a, double dx, double dy, double dz, double gx, double gy,
                                                  double gz, double A, double B3) {
    uj_accumulate(a, dx, dy, dz, gx, gy, gz, A, B3);
}
__device__ __forceinline__ void acc_add_far(UJAcc& t, const UJAcc& a) {
    t.u0 += a.u0; t.u1 += a.u1; t.u2 += a.u2;
    t.j0 = fma(-3.0, a.j0, t.j0); t.j1 = fma(-3.0, a.j1, t.j1); t.j2 = fma(-3.0, a.j2, t.j2);
    t.j3 = fma(-3.0, a.j3, t.j3); t.j4 = fma(-3.0, a.j4, t.j4); t.j5 = fma(-3.0, a.j5, t.j5);
    t.j6 = fma(-3.0, a.j6, t.j6); t.j7 = fma(-3.0, a.j7, t.j7);
    t.w0 += a.w0; t.w1 += a.w1; t.w2 += a.w2;
}
// j8 <- trace0 - (j0 + j4): closes the trace after the pair loops (trace0 = trace of whatever J was put into the
// accumulators before the loops: 0 for the direct kernels, the L2P part for the FMM leaf kernel).
__device__ __forceinline__ void acc_close_trace(UJAcc& a, double trace0 = 0.0) { a.j8 = trace0 - (a.j0 + a.j4); }

// x != 0 for a non-negative double, on the integer pipe (keeps DSETP off the FP64 pipe)
__device__ __forceinline__ bool nonzero_f64(double x) { return (__double2hiint(x) | __double2loint(x)) != 0; }

// Far field / singular kernel: A = 1/r^3, B = -3/r^5.  Requires r2 > 0.
__device__ __forceinline__ void ab_singular(double r2, double& A, double& B) {
    double ri = rsqrt_f64(r2);
    double ri2 = ri * ri;
    A = ri2 * ri;
    B = (-3.0 * ri2) * A;
}

// Gaussian-erf near field from the table: t < T_FAR.  H = 2 dG/dt, so one degree-9 polynomial per interval gives both:
// value and derivative share the Horner recurrence (17 FMAs) and a lookup reads 80 B (five LDS.128) instead of two
// polynomials' 128 B — the pair loops are co-limited by shared-memory wavefronts and the FP64 pipe (tools/gen_tables.py).
// REP > 1: the table is stored REP times, entry (j, i) of copy q at tab[(j * NINT + i) * REP + q]; with REP = 8 and
// q = lane & 7 every lane of a quarter-warp reads its own 16-byte bank group, so a lookup is bank-conflict free whatever
// the lanes' interval indices are (4 wavefronts per LDS.128 instead of the measured 6.2).
template <int REP>
__device__ __forceinline__ void ab_gauss_table_rep(const double2* __restrict__ tab, int q, double t, double sinv3, double sinv5,
                                                   double& A, double& B) {
    static_assert(VPM_GG_DEG % 2 == 1, "coefficients are stored in pairs");
    double m = fma(t, VPM_GG_INVW, MAGIC_RINT);
    int i = __double2loint(m);
    double u = fma(m - MAGIC_RINT, -VPM_GG_W, t);
    const double2* tp = tab + (REP == 1 ? i : i * REP + q);
    double2 c[(VPM_GG_DEG + 1) / 2];
#pragma unroll
    for (int j = 0; j < (VPM_GG_DEG + 1) / 2; ++j) c[j] = tp[j * VPM_GG_NINT * REP];   // {c_2j, c_2j+1}
    double d = c[(VPM_GG_DEG - 1) / 2].y;   // derivative runs one step behind the value
    double p = fma(d, u, c[(VPM_GG_DEG - 1) / 2].x);
#pragma unroll
    for (int j = (VPM_GG_DEG - 1) / 2 - 1; j >= 0; --j) {
        d = fma(d, u, p); p = fma(p, u, c[j].y);
        d = fma(d, u, p); p = fma(p, u, c[j].x);
    }
    A = p * sinv3;
    B = (d + d) * sinv5;
}
__device__ __forceinline__ void ab_gauss_table(const double2* __restrict__ tab, double t, double sinv3, double sinv5,
                                               double& A, double& B) {
    ab_gauss_table_rep<1>(tab, 0, t, sinv3, sinv5, A, B);
}

// Winckelmans: G = (t + 2.5)/(t+1)^2.5, H = -(3 t + 10.5)/(t+1)^3.5   (SURVEY.md A.3 in the variable t)
__device__ __forceinline__ void ab_winckelmans(double t, double sinv3, double sinv5, double& A, double& B) {
    double w = rsqrt_f64(t + 1.0);
    double w2 = w * w;
    double w4 = w2 * w2;
    double w5 = w4 * w;
    double w7 = w5 * w2;
    A = ((t + 2.5) * w5) * sinv3;
    B = (fma(-3.0, t, -10.5) * w7) * sinv5;
}

// exp(x) and 1 - exp(x) for x <= 0 in FP64 (|rel err| < 3e-16 each): Cody-Waite reduction + degree-12 Horner.
// 1 - exp(x) is formed from the polynomial without its leading 1 when no scaling is needed, so it keeps full
// relative accuracy as x -> 0 (the reference's `1 - exp(-r^3)` cancels there; we are more accurate, not less).
__device__ __forceinline__ void exp_neg_f64(double x, double& e, double& one_minus_e) {
    if (x < -708.0) { e = 0.0; one_minus_e = 1.0; return; }
    const double L2E = 1.4426950408889634, LN2H = 6.93147180369123816490e-01, LN2L = 1.90821492927058770002e-10;
    double m = fma(x, L2E, MAGIC_RINT);
    int k = __double2loint(m);
    double kf = m - MAGIC_RINT;
    double r = fma(kf, -LN2H, x);
    r = fma(kf, -LN2L, r);
    double p = 2.08767569878680989792e-09;                 // 1/12!
    p = fma(p, r, 2.50521083854417187751e-08);             // 1/11!
    p = fma(p, r, 2.75573192239858906526e-07);             // 1/10!
    p = fma(p, r, 2.75573192239858906526e-06);             // 1/9!
    p = fma(p, r, 2.48015873015873015873e-05);             // 1/8!
    p = fma(p, r, 1.98412698412698412698e-04);             // 1/7!
    p = fma(p, r, 1.38888888888888888889e-03);             // 1/6!
    p = fma(p, r, 8.33333333333333333333e-03);             // 1/5!
    p = fma(p, r, 4.16666666666666666667e-02);             // 1/4!
    p = fma(p, r, 1.66666666666666666667e-01);             // 1/3!
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    double q = p * r;                                      // exp(r) - 1
    if (k == 0) {
        e = 1.0 + q;
        one_minus_e = -q;
    } else {
        double er = 1.0 + q;
        // scale by 2^k through the exponent field (k in [-1022, -1])
        e = __hiloint2double(__double2hiint(er) + k * 1048576, __double2loint(er));
        one_minus_e = 1.0 - e;
    }
}

// `gaussian` kernel: g = 1 - exp(-s^3), g' = 3 s^2 exp(-s^3).  Requires t > 0.
__device__ __forceinline__ void ab_gaussian(double t, double sinv3, double sinv5, double& A, double& B) {
    double rs = rsqrt_f64(t);       // 1/s
    double s3 = t * (t * rs);       // s^3
    double e, ome;
    exp_neg_f64(-s3, e, ome);
    double rs2 = rs * rs;
    double G = ome * (rs2 * rs);
    double H = (3.0 * (e - G)) * rs2;
    A = G * sinv3;
    B = H * sinv5;
}

// ---- record accessors (layout in common.cuh) ---------------------------------------------------------------
struct SrcCore {  // what every branch needs: three LDS.128
    double x, y, z, gx, gy, gz;
};
__device__ __forceinline__ SrcCore load_core(const double2* __restrict__ rec) {
    const double2 q0 = rec[0], q1 = rec[1], q2 = rec[2];
    return SrcCore{q0.x, q0.y, q1.x, q1.y, q2.x, q2.y};
}

// Far-field interaction (also the whole `singular` kernel when guarded): no table, no branch.
__device__ __forceinline__ void uj_pair_far(UJAcc& a, double tx, double ty, double tz, const SrcCore& s) {
    double dx = tx - s.x, dy = ty - s.y, dz = tz - s.z;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double ri = rsqrt_f64(r2);
    double ri2 = ri * ri;
    double A = ri2 * ri;
    uj_accumulate_far(a, dx, dy, dz, s.gx, s.gy, s.gz, A, ri2 * A);   // B / (-3); the caller adds with acc_add_far
}

// General interaction of the regularised kernels: per-lane choice between table/closed form and far field.
template <int KERNEL, int REP = 1>
__device__ __forceinline__ void uj_pair_general(UJAcc& a, double dx, double dy, double dz, double r2, const SrcCore& s,
                                                const double2* __restrict__ rec, const double2* __restrict__ tab, int q = 0) {
    const double2 q3 = rec[3];  // T_FAR sigma^2, 1/sigma^3
    const double2 q4 = rec[4];  // 1/sigma^5, 1/sigma^2
    double A, B;
    if (KERNEL == K_GAUSSIANERF) {
        if (__double2hiint(r2) > __double2hiint(q3.x)) {
            ab_singular(r2, A, B);
        } else {
            ab_gauss_table_rep<REP>(tab, q, r2 * q4.y, q3.y, q4.x, A, B);
            A = nonzero_f64(r2) ? A : 0.0;  // r == 0 skip (src/FLOWUnsteady_processing_force.jl:895)
        }
    } else if (KERNEL == K_WINCKELMANS) {
        ab_winckelmans(r2 * q4.y, q3.y, q4.x, A, B);
        A = nonzero_f64(r2) ? A : 0.0;
    } else {  // K_GAUSSIAN
        const double t = r2 * q4.y;
        const bool nz = nonzero_f64(t);
        ab_gaussian(nz ? t : 1.0, q3.y, q4.x, A, B);
        A = nz ? A : 0.0;
        B = nz ? B : 0.0;
    }
    uj_accumulate(a, dx, dy, dz, s.gx, s.gy, s.gz, A, B);
}

// Two sources at a time.  For gaussianerf a warp-uniform vote on "both sources are far for all 32 lanes" selects a
// straight-line block with two independent rsqrt chains (ILP for the 2-cycle-issue FP64 pipe); the far test is an
// integer compare of high words (for positive doubles hi(a) > hi(b) implies a > b; conservative by < 2^-20, which the
// table's spare intervals cover), so it costs no FP64 issue slot.
template <int KERNEL>
__device__ __forceinline__ void uj_pair2(UJAcc& a, double tx, double ty, double tz, const double2* __restrict__ rec,
                                         const double2* __restrict__ tab) {
    const double2* recA = rec;
    const double2* recB = rec + REC_REALS / 2;
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
        const bool far = __double2hiint(r2a) > __double2hiint(recA[3].x) && __double2hiint(r2b) > __double2hiint(recB[3].x);
        if (__all_sync(0xffffffffu, far)) {
            double Aa, Ba, Ab, Bb;
            ab_singular(r2a, Aa, Ba);
            ab_singular(r2b, Ab, Bb);
            uj_accumulate(a, dxa, dya, dza, sa.gx, sa.gy, sa.gz, Aa, Ba);
            uj_accumulate(a, dxb, dyb, dzb, sb.gx, sb.gy, sb.gz, Ab, Bb);
        } else {
            uj_pair_general<KERNEL>(a, dxa, dya, dza, r2a, sa, recA, tab);
            uj_pair_general<KERNEL>(a, dxb, dyb, dzb, r2b, sb, recB, tab);
        }
    } else {
        uj_pair_general<KERNEL>(a, dxa, dya, dza, r2a, sa, recA, tab);
        uj_pair_general<KERNEL>(a, dxb, dyb, dzb, r2b, sb, recB, tab);
    }
}

// Shared-memory layout of the pairwise kernels.
struct __align__(16) PairSmem {
    double tile[2][TILE_DOUBLES];
    uint64_t full[2];
    double tbox[8];  // bounding box of this CTA's targets: xmin, xmax, ymin, ymax, zmin, zmax
    double red[6][8];
};

constexpr int UJ_BT = 256;  // targets (= threads) per CTA
constexpr uint32_t TILE_BYTES = TILE_DOUBLES * sizeof(double);

// Bounding box of the CTA's live targets -> sm.tbox (all threads must call).
__device__ __forceinline__ void cta_target_box(PairSmem& sm, bool live, double px, double py, double pz) {
    const double big = 1.0e300;
    double v[6] = {live ? px : big, live ? -px : big, live ? py : big, live ? -py : big, live ? pz : big, live ? -pz : big};
#pragma unroll
    for (int c = 0; c < 6; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[c] = fmin(v[c], __shfl_xor_sync(0xffffffffu, v[c], o));
    }
    const int w = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0)
        for (int c = 0; c < 6; ++c) sm.red[c][w] = v[c];
    __syncthreads();
    if (threadIdx.x < 6) {
        double m = sm.red[threadIdx.x][0];
        for (int k = 1; k < UJ_BT / 32; ++k) m = fmin(m, sm.red[threadIdx.x][k]);
        sm.tbox[threadIdx.x] = (threadIdx.x & 1) ? -m : m;  // odd slots hold maxima
    }
    __syncthreads();
}

// Squared distance between the CTA's target box and a tile's source box (0 when they overlap).
__device__ __forceinline__ double box_dist2(const double* __restrict__ tb, const double* __restrict__ hdr) {
    double d2 = 0.0;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        double gap = fmax(fmax(tb[2 * c] - hdr[2 * c + 1], hdr[2 * c] - tb[2 * c + 1]), 0.0);
        d2 = fma(gap, gap, d2);
    }
    return d2;
}

constexpr size_t uj_smem_bytes(int kernel) {
    return sizeof(PairSmem) + (kernel == K_GAUSSIANERF ? sizeof(double) * VPM_GG_DOUBLES : 0);
}

// Source-split launch geometry: when there are too few target blocks to fill the GPU for many waves (small N, or a
// rank's shard of a multi-GPU run) the source tiles are split into `nchunks` ranges and the grid becomes
// (target blocks) x (chunks); chunk c writes its partial U, J into rows [12 c, 12 c + 12) of `partial` (row length ldp)
// and reduce_partials_kernel sums the chunks in a fixed order (deterministic, unlike atomics).
struct SplitArgs {
    double* partial;       // nullptr: single chunk, write U/J directly
    int64_t ldp;
    int tiles_per_chunk;
};

// Grid: (ceil(nt / UJ_BT), nchunks) CTAs.  srec: ntiles tiles of TILE_DOUBLES doubles (records + header, common.cuh).
// Targets: positions tx/ty/tz (nt each).  Outputs: component k of U at U[k * ldo + i], of J at J[k * ldo + i].
template <int KERNEL>
__global__ void __launch_bounds__(UJ_BT, 2)
uj_direct_f64_kernel(const double* __restrict__ srec, int ntiles, const double* __restrict__ tx,
                     const double* __restrict__ ty, const double* __restrict__ tz, int64_t nt, double* __restrict__ U,
                     double* __restrict__ J, int64_t ldo, int accumulate, const double* __restrict__ gh_table,
                     SplitArgs split) {
    if (split.partial != nullptr) {
        const int t0 = blockIdx.y * split.tiles_per_chunk;
        srec += (size_t)t0 * TILE_DOUBLES;
        ntiles = min(ntiles - t0, split.tiles_per_chunk);
        U = split.partial + (size_t)blockIdx.y * 12 * split.ldp;
        J = U + 3 * split.ldp;
        ldo = split.ldp;
        accumulate = 0;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>(smem_raw);
    double2* tab = reinterpret_cast<double2*>(smem_raw + sizeof(PairSmem));

    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * UJ_BT + tid;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        fence_mbar_init();
    }
    if (KERNEL == K_GAUSSIANERF) {
        const double2* g2 = reinterpret_cast<const double2*>(gh_table);
        for (int k = tid; k < VPM_GG_DOUBLES / 2; k += UJ_BT) tab[k] = g2[k];
    }
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(&sm.full[0], TILE_BYTES);
        bulk_g2s(sm.tile[0], srec, TILE_BYTES, &sm.full[0]);
    }

    const bool live = i < nt;
    const double px = live ? tx[i] : 0.0, py = live ? ty[i] : 0.0, pz = live ? tz[i] : 0.0;
    if (KERNEL == K_GAUSSIANERF) cta_target_box(sm, live, px, py, pz);
    UJAcc tot;
    acc_zero(tot);

    for (int k = 0; k < ntiles; ++k) {
        const int b = k & 1;
        if (tid == 0 && k + 1 < ntiles) {
            // buffer b^1 was last read in iteration k-1; the __syncthreads() that closed it orders those reads
            mbar_arrive_expect_tx(&sm.full[b ^ 1], TILE_BYTES);
            bulk_g2s(sm.tile[b ^ 1], srec + (size_t)(k + 1) * TILE_DOUBLES, TILE_BYTES, &sm.full[b ^ 1]);
        }
        mbar_wait(&sm.full[b], (k >> 1) & 1);
        const double2* rec = reinterpret_cast<const double2*>(sm.tile[b]);
        UJAcc a;
        acc_zero(a);
        bool tile_far = false;
        if (KERNEL == K_GAUSSIANERF) {
            const double* hdr = sm.tile[b] + TILE_HDR;
            tile_far = box_dist2(sm.tbox, hdr) > hdr[6];  // CTA-uniform: whole tile beyond T_FAR sigma_max^2
        }
        if (tile_far) {
            // every (target, source) pair of this tile is in the singular regime: no vote, no table, full ILP
#pragma unroll 4
            for (int j = 0; j < TILE_SRC; ++j) uj_pair_far(a, px, py, pz, load_core(rec + j * (REC_REALS / 2)));
            acc_add_far(tot, a);
        } else {
#pragma unroll 1
            for (int j = 0; j < TILE_SRC; j += 2) uj_pair2<KERNEL>(a, px, py, pz, rec + j * (REC_REALS / 2), tab);
            acc_add(tot, a);
        }
        __syncthreads();
    }

    if (live) {
        acc_close_trace(tot);
        // expand the delta term: J[2,1] -= w3, J[3,1] += w2, J[1,2] += w3, J[3,2] -= w1, J[1,3] -= w2, J[2,3] += w1
        tot.j1 -= tot.w2; tot.j2 += tot.w1;
        tot.j3 += tot.w2; tot.j5 -= tot.w0;
        tot.j6 -= tot.w1; tot.j7 += tot.w0;
        double o[12] = {tot.u0, tot.u1, tot.u2, tot.j0, tot.j1, tot.j2, tot.j3, tot.j4, tot.j5, tot.j6, tot.j7, tot.j8};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double* p = U + (size_t)c * ldo + i;
            *p = accumulate ? *p + o[c] : o[c];
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            double* p = J + (size_t)c * ldo + i;
            *p = accumulate ? *p + o[3 + c] : o[3 + c];
        }
    }
}

}  // namespace vpm
