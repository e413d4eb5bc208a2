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
//     (tools/gen_tables.py, max rel. error 3e-15).  For t >= T_FAR the Gaussian-erf kernel equals the
//     singular kernel to < 2^-56, and a warp-uniform vote takes the branch with one MUFU.RSQ64H + 5 DFMA;
//   * -1/(4 pi) is folded into the source's Gamma, the antisymmetric (Kronecker-delta) part of J is
//     accumulated as a 3-vector and expanded once at the end, and per-tile partial sums are added to the
//     running totals (pairwise-style summation keeps the 1e6-term sums inside the 1e-12 parity budget).
#pragma once

#include "common.cuh"
#include "gauss_table.inc"

namespace vpm {

constexpr double MAGIC_RINT = 6755399441055744.0;  // 1.5 * 2^52: fma(t, 1/W, MAGIC) puts rint(t/W) in the low word

struct UJAcc {
    double u0, u1, u2;
    double j0, j1, j2, j3, j4, j5, j6, j7, j8;  // J[i + 3 j]
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
    t.j5 += a.j5; t.j6 += a.j6; t.j7 += a.j7; t.j8 += a.j8;
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
    a.j6 = fma(b0, dz, a.j6); a.j7 = fma(b1, dz, a.j7); a.j8 = fma(b2, dz, a.j8);
    a.w0 = fma(A, gx, a.w0);
    a.w1 = fma(A, gy, a.w1);
    a.w2 = fma(A, gz, a.w2);
}

// Far field / singular kernel: A = 1/r^3, B = -3/r^5.  Requires r2 > 0.
__device__ __forceinline__ void ab_singular(double r2, double& A, double& B) {
    double ri = rsqrt_f64(r2);
    double ri2 = ri * ri;
    A = ri2 * ri;
    B = (-3.0 * ri2) * A;
}

// Gaussian-erf near field from the table: t < T_FAR.
__device__ __forceinline__ void ab_gauss_table(const double2* __restrict__ tab, double t, double sinv3, double sinv5,
                                               double& A, double& B) {
    double m = fma(t, VPM_GT_INVW, MAGIC_RINT);
    int i = __double2loint(m);
    double u = fma(m - MAGIC_RINT, -VPM_GT_W, t);
    const double2* tp = tab + i;
    double2 c = tp[VPM_GT_DEG * VPM_GT_NINT];
    double G = c.x, H = c.y;
#pragma unroll
    for (int k = VPM_GT_DEG - 1; k >= 0; --k) {
        c = tp[k * VPM_GT_NINT];
        G = fma(G, u, c.x);
        H = fma(H, u, c.y);
    }
    A = G * sinv3;
    B = H * sinv5;
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

// exp(x) for x <= 0 in FP64 (|rel err| < 2e-16): Cody-Waite reduction + degree-11 Taylor/Horner.
__device__ __forceinline__ double exp_neg_f64(double x) {
    if (x < -708.0) return 0.0;
    const double L2E = 1.4426950408889634, LN2H = 6.93147180369123816490e-01, LN2L = 1.90821492927058770002e-10;
    double m = fma(x, L2E, MAGIC_RINT);
    int k = __double2loint(m);
    double kf = m - MAGIC_RINT;
    double r = fma(kf, -LN2H, x);
    r = fma(kf, -LN2L, r);
    double p = 2.50521083854417187751e-08;                 // 1/11!
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
    p = fma(p, r, 1.0);
    // scale by 2^k through the exponent field (k in [-1022, 0])
    return __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
}

// `gaussian` kernel: g = 1 - exp(-s^3), g' = 3 s^2 exp(-s^3).  Requires t > 0.
__device__ __forceinline__ void ab_gaussian(double t, double sinv3, double sinv5, double& A, double& B) {
    double rs = rsqrt_f64(t);       // 1/s
    double s3 = t * (t * rs);       // s^3
    double e = exp_neg_f64(-s3);
    double rs2 = rs * rs;
    double G = (1.0 - e) * (rs2 * rs);
    double H = (3.0 * (e - G)) * rs2;
    A = G * sinv3;
    B = H * sinv5;
}

// One (target, source) interaction.  `rec` points at the 10-double source record in shared memory.
template <int KERNEL>
__device__ __forceinline__ void uj_pair(UJAcc& a, double tx, double ty, double tz, const double2* __restrict__ rec,
                                        const double2* __restrict__ tab) {
    const double2 s0 = rec[0];  // x, y
    const double2 s1 = rec[1];  // z, 1/sigma^2
    const double2 s2 = rec[2];  // G'x, G'y
    const double2 s3 = rec[3];  // G'z, 1/sigma^3
    double dx = tx - s0.x, dy = ty - s0.y, dz = tz - s1.x;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double A, B;
    if (KERNEL == K_SINGULAR) {
        ab_singular(r2 > 0.0 ? r2 : 1.0, A, B);
        A = r2 > 0.0 ? A : 0.0;  // r == 0 skip (src/FLOWUnsteady_processing_force.jl:895)
    } else if (KERNEL == K_GAUSSIANERF) {
        double t = r2 * s1.y;
        if (__all_sync(0xffffffffu, t >= VPM_GT_TFAR)) {
            ab_singular(r2, A, B);  // t >= T_FAR > 0 implies r2 > 0
        } else {
            if (t < VPM_GT_TFAR) {
                ab_gauss_table(tab, t, s3.y, rec[4].x, A, B);
                A = r2 > 0.0 ? A : 0.0;
            } else {
                ab_singular(r2, A, B);
            }
        }
    } else if (KERNEL == K_WINCKELMANS) {
        double t = r2 * s1.y;
        ab_winckelmans(t, s3.y, rec[4].x, A, B);
        A = r2 > 0.0 ? A : 0.0;
    } else {  // K_GAUSSIAN
        double t = r2 * s1.y;
        ab_gaussian(t > 0.0 ? t : 1.0, s3.y, rec[4].x, A, B);
        A = r2 > 0.0 ? A : 0.0;
        B = r2 > 0.0 ? B : 0.0;
    }
    uj_accumulate(a, dx, dy, dz, s2.x, s2.y, s3.x, A, B);
}

// Shared-memory layout of the pairwise kernels.
struct __align__(16) PairSmem {
    double tile[2][TILE_SRC * REC_REALS];
    uint64_t full[2];
};

constexpr int UJ_BT = 256;  // targets (= threads) per CTA

constexpr size_t uj_smem_bytes(int kernel) {
    return sizeof(PairSmem) + (kernel == K_GAUSSIANERF ? sizeof(double) * 2 * (VPM_GT_DEG + 1) * VPM_GT_NINT : 0);
}

// Grid: ceil(nt / UJ_BT) CTAs.  srec: ntiles * TILE_SRC records (tail padded with null records).
// Targets: positions tx/ty/tz (nt each).  Outputs: component k of U at U[k * ldo + i], of J at J[k * ldo + i].
template <int KERNEL>
__global__ void __launch_bounds__(UJ_BT, 2)
uj_direct_f64_kernel(const double* __restrict__ srec, int ntiles, const double* __restrict__ tx,
                     const double* __restrict__ ty, const double* __restrict__ tz, int64_t nt, double* __restrict__ U,
                     double* __restrict__ J, int64_t ldo, int accumulate, const double* __restrict__ gh_table) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>(smem_raw);
    double2* tab = reinterpret_cast<double2*>(smem_raw + sizeof(PairSmem));

    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * UJ_BT + tid;
    constexpr uint32_t TILE_BYTES = TILE_SRC * REC_REALS * sizeof(double);

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        fence_mbar_init();
    }
    if (KERNEL == K_GAUSSIANERF) {
        const double2* g2 = reinterpret_cast<const double2*>(gh_table);
        for (int k = tid; k < (VPM_GT_DEG + 1) * VPM_GT_NINT; k += UJ_BT) tab[k] = g2[k];
    }
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(&sm.full[0], TILE_BYTES);
        bulk_g2s(sm.tile[0], srec, TILE_BYTES, &sm.full[0]);
    }

    const bool live = i < nt;
    const double px = live ? tx[i] : 0.0, py = live ? ty[i] : 0.0, pz = live ? tz[i] : 0.0;
    UJAcc tot;
    acc_zero(tot);

    for (int k = 0; k < ntiles; ++k) {
        const int b = k & 1;
        if (tid == 0 && k + 1 < ntiles) {
            // buffer b^1 was last read in iteration k-1; the __syncthreads() that closed it orders those reads
            mbar_arrive_expect_tx(&sm.full[b ^ 1], TILE_BYTES);
            bulk_g2s(sm.tile[b ^ 1], srec + (size_t)(k + 1) * TILE_SRC * REC_REALS, TILE_BYTES, &sm.full[b ^ 1]);
        }
        mbar_wait(&sm.full[b], (k >> 1) & 1);
        const double2* rec = reinterpret_cast<const double2*>(sm.tile[b]);
        UJAcc a;
        acc_zero(a);
#pragma unroll 2
        for (int j = 0; j < TILE_SRC; ++j) uj_pair<KERNEL>(a, px, py, pz, rec + j * (REC_REALS / 2), tab);
        acc_add(tot, a);
        __syncthreads();
    }

    if (live) {
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
