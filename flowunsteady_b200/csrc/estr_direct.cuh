// estr_direct.cuh — K2: pairwise SFS vortex-stretching term E_str (FLOWVPM's Estr_direct / the near-field
// part of Estr_fmm; docs at /root/reference/docs/src/api/flowvpm-sfs.md:12-13, model equation at
// /root/reference/docs/src/theory/rvpm.md:251-263, index form SURVEY.md A.4):
//
//     SFS_p += sum_q zeta_sigma_q(x_p - x_q) * S(J_p - J_q, Gamma_q),   S_k = sum_l dJ[l,k] Gamma_l (transposed)
//
// Because S is linear in J, the sum splits into  J_p^T a_p - b_p  with
//     a_p = sum_q zeta_q Gamma_q      (3 accumulators),   b_p = sum_q zeta_q (J_q^T Gamma_q)   (3 accumulators)
// so the pair loop never touches J_p and the source carries v_q = J_q^T Gamma_q precomputed (same 80-byte record
// size as the UJ pass).  zeta is evaluated in t = r^2/sigma^2 from the shared-memory table Z(t) = exp(-t/2);
// beyond T_FAR, zeta/zeta(0) < 8e-20 and a warp-uniform vote skips the pair.
#pragma once

#include "uj_direct.cuh"

namespace vpm {

struct EAcc {
    double a0, a1, a2, b0, b1, b2;
};

// Z(t) = exp(-t/2) = E[i] * exp(-u/2), i = rint(t / WZ), u = t - i WZ: one tabulated double per lookup, the second
// factor is a degree-7 Taylor polynomial with immediate coefficients (|u/2| <= 1/32).  Requires 0 <= t < T_FAR.
__device__ __forceinline__ double zeta_gauss_table(const double* __restrict__ ztab, double t) {
    double m = fma(t, VPM_GZ_INVW, MAGIC_RINT);
    int i = __double2loint(m);
    double u = fma(m - MAGIC_RINT, -VPM_GZ_W, t);
    const double e = ztab[i];
    double z = VPM_GZ_C7;
    z = fma(z, u, VPM_GZ_C6);
    z = fma(z, u, VPM_GZ_C5);
    z = fma(z, u, VPM_GZ_C4);
    z = fma(z, u, VPM_GZ_C3);
    z = fma(z, u, VPM_GZ_C2);
    z = fma(z, u, VPM_GZ_C1);
    z = fma(z, u, VPM_GZ_C0);
    return z * e;
}

template <int KERNEL>
__device__ __forceinline__ void estr_pair(EAcc& a, double tx, double ty, double tz, const double2* __restrict__ rec,
                                          const double* __restrict__ ztab) {
    const double2 s0 = rec[0];  // x, y
    const double2 s1 = rec[1];  // z, 1/sigma^2
    double dx = tx - s0.x, dy = ty - s0.y, dz = tz - s1.x;
    double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
    double t = r2 * s1.y;
    double z;
    if (KERNEL == K_GAUSSIANERF) {
        if (__all_sync(0xffffffffu, t >= VPM_GT_TFAR)) return;
        z = t < VPM_GT_TFAR ? zeta_gauss_table(ztab, t < VPM_GT_TFAR ? t : 0.0) : 0.0;
    } else if (KERNEL == K_WINCKELMANS) {
        double w = rsqrt_f64(t + 1.0);
        double w2 = w * w;
        z = (w2 * w2) * (w2 * w);  // (t+1)^-3.5
    } else if (KERNEL == K_GAUSSIAN) {
        double s3 = t > 0.0 ? t * (t * rsqrt_f64(t)) : 0.0;
        double ome;
        exp_neg_f64(-s3, z, ome);
    } else {
        z = r2 == 0.0 ? 1.0 : 0.0;
    }
    const double2 s2 = rec[2];  // cGx, cGy
    const double2 s3 = rec[3];  // cGz, cvx
    const double2 s4 = rec[4];  // cvy, cvz
    a.a0 = fma(z, s2.x, a.a0);
    a.a1 = fma(z, s2.y, a.a1);
    a.a2 = fma(z, s3.x, a.a2);
    a.b0 = fma(z, s3.y, a.b0);
    a.b1 = fma(z, s4.x, a.b1);
    a.b2 = fma(z, s4.y, a.b2);
}

constexpr size_t estr_smem_bytes(int kernel) {
    return sizeof(PairSmem) + (kernel == K_GAUSSIANERF ? sizeof(double) * VPM_GZ_NINT : 0);
}

// Targets: positions tx/ty/tz and their J (component c at Jt[c * ldj + i]).  SFS component k accumulates at
// SFS[k * ldo + i] (callers reset it first when `reset_sfs` is requested).
template <int KERNEL>
__global__ void __launch_bounds__(UJ_BT, 2)
estr_direct_f64_kernel(const double* __restrict__ srec, int ntiles, const double* __restrict__ tx,
                       const double* __restrict__ ty, const double* __restrict__ tz, int64_t nt,
                       const double* __restrict__ Jt, int64_t ldj, int transposed, double* __restrict__ SFS,
                       int64_t ldo, const double* __restrict__ z_table, SplitArgs split, int accumulate, int raw) {
    if (split.partial != nullptr) {
        const int t0 = blockIdx.y * split.tiles_per_chunk;
        srec += (size_t)t0 * TILE_DOUBLES;
        ntiles = min(ntiles - t0, split.tiles_per_chunk);
        SFS = split.partial + (size_t)blockIdx.y * 3 * split.ldp;
        ldo = split.ldp;
        accumulate = 0;
    }
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PairSmem& sm = *reinterpret_cast<PairSmem*>(smem_raw);
    double* ztab = reinterpret_cast<double*>(smem_raw + sizeof(PairSmem));

    const int tid = threadIdx.x;
    const int64_t i = (int64_t)blockIdx.x * UJ_BT + tid;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        fence_mbar_init();
    }
    if (KERNEL == K_GAUSSIANERF)
        for (int k = tid; k < VPM_GZ_NINT; k += UJ_BT) ztab[k] = z_table[k];
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(&sm.full[0], TILE_BYTES);
        bulk_g2s(sm.tile[0], srec, TILE_BYTES, &sm.full[0]);
    }

    const bool live = i < nt;
    const double px = live ? tx[i] : 0.0, py = live ? ty[i] : 0.0, pz = live ? tz[i] : 0.0;
    cta_target_box(sm, live, px, py, pz);
    EAcc tot = {0, 0, 0, 0, 0, 0};

    for (int k = 0; k < ntiles; ++k) {
        const int b = k & 1;
        if (tid == 0 && k + 1 < ntiles) {
            mbar_arrive_expect_tx(&sm.full[b ^ 1], TILE_BYTES);
            bulk_g2s(sm.tile[b ^ 1], srec + (size_t)(k + 1) * TILE_DOUBLES, TILE_BYTES, &sm.full[b ^ 1]);
        }
        mbar_wait(&sm.full[b], (k >> 1) & 1);
        const double2* rec = reinterpret_cast<const double2*>(sm.tile[b]);
        const double* hdr = sm.tile[b] + TILE_HDR;
        // CTA-uniform: a tile whose every source is beyond T_FAR sigma^2 of every target contributes zeta < 8e-20 zeta(0)
        if (!(box_dist2(sm.tbox, hdr) > hdr[6])) {
            EAcc a = {0, 0, 0, 0, 0, 0};
#pragma unroll 2
            for (int j = 0; j < TILE_SRC; ++j) estr_pair<KERNEL>(a, px, py, pz, rec + j * (REC_REALS / 2), ztab);
            tot.a0 += a.a0; tot.a1 += a.a1; tot.a2 += a.a2;
            tot.b0 += a.b0; tot.b1 += a.b1; tot.b2 += a.b2;
        }
        __syncthreads();
    }

    if (live && raw) {
        // zeta pass (vpm.zeta_direct): the plain kernel sum  sum_q zeta_sigma_q(x_p - x_q) v_q  (records carry c v in the
        // Gamma slots); used for the vorticity and as the matrix-vector product of the RBF conjugate gradient
        SFS[0 * ldo + i] = (accumulate ? SFS[0 * ldo + i] : 0.0) + tot.a0;
        SFS[1 * ldo + i] = (accumulate ? SFS[1 * ldo + i] : 0.0) + tot.a1;
        SFS[2 * ldo + i] = (accumulate ? SFS[2 * ldo + i] : 0.0) + tot.a2;
    } else if (live) {
        double Jp[9];
#pragma unroll
        for (int c = 0; c < 9; ++c) Jp[c] = Jt[(size_t)c * ldj + i];
        double e0, e1, e2;
        if (transposed) {  // S_k = sum_l J[l,k] a_l ; J[l,k] at l + 3 k
            e0 = Jp[0] * tot.a0 + Jp[1] * tot.a1 + Jp[2] * tot.a2;
            e1 = Jp[3] * tot.a0 + Jp[4] * tot.a1 + Jp[5] * tot.a2;
            e2 = Jp[6] * tot.a0 + Jp[7] * tot.a1 + Jp[8] * tot.a2;
        } else {           // S_k = sum_l J[k,l] a_l
            e0 = Jp[0] * tot.a0 + Jp[3] * tot.a1 + Jp[6] * tot.a2;
            e1 = Jp[1] * tot.a0 + Jp[4] * tot.a1 + Jp[7] * tot.a2;
            e2 = Jp[2] * tot.a0 + Jp[5] * tot.a1 + Jp[8] * tot.a2;
        }
        SFS[0 * ldo + i] = (accumulate ? SFS[0 * ldo + i] : 0.0) + (e0 - tot.b0);
        SFS[1 * ldo + i] = (accumulate ? SFS[1 * ldo + i] : 0.0) + (e1 - tot.b1);
        SFS[2 * ldo + i] = (accumulate ? SFS[2 * ldo + i] : 0.0) + (e2 - tot.b2);
    }
}

}  // namespace vpm
