// uj_direct_f32.cuh — FP32 variant of K1 (the reference's knob is `vpm_floattype`,
// /root/reference/src/FLOWUnsteady_simulation.jl:137).  Pair arithmetic and per-tile partial sums are FP32
// (128 FFMA lanes/SM vs 64 DFMA lanes/SM... i.e. the FP32 pipe, 2x the issue rate and no MUFU.RSQ64H fix-up);
// the state, the source records in HBM and the cross-tile totals stay FP64.  Records are converted to FP32 once per
// tile when they are staged, relative to the CTA's first target so close pairs keep their separation digits.
#pragma once

#include "uj_direct.cuh"

namespace vpm {

struct UJAcc32 {
    float u0, u1, u2, j0, j1, j2, j3, j4, j5, j6, j7, j8, w0, w1, w2;
};

__device__ __forceinline__ void uj_accumulate32(UJAcc32& a, float dx, float dy, float dz, float gx, float gy, float gz,
                                                float A, float B) {
    float c0 = fmaf(dy, gz, -dz * gy);
    float c1 = fmaf(dz, gx, -dx * gz);
    float c2 = fmaf(dx, gy, -dy * gx);
    a.u0 = fmaf(A, c0, a.u0);
    a.u1 = fmaf(A, c1, a.u1);
    a.u2 = fmaf(A, c2, a.u2);
    float b0 = B * c0, b1 = B * c1, b2 = B * c2;
    a.j0 = fmaf(b0, dx, a.j0); a.j1 = fmaf(b1, dx, a.j1); a.j2 = fmaf(b2, dx, a.j2);
    a.j3 = fmaf(b0, dy, a.j3); a.j4 = fmaf(b1, dy, a.j4); a.j5 = fmaf(b2, dy, a.j5);
    a.j6 = fmaf(b0, dz, a.j6); a.j7 = fmaf(b1, dz, a.j7); a.j8 = fmaf(b2, dz, a.j8);
    a.w0 = fmaf(A, gx, a.w0);
    a.w1 = fmaf(A, gy, a.w1);
    a.w2 = fmaf(A, gz, a.w2);
}

// FP32 tile record: { dx0, dy0, dz0, 1/sigma^2 | G'x, G'y, G'z, 1/sigma^3 | 1/sigma^5 } with positions relative to
// the CTA origin; 12 floats (48 B) so it moves as three LDS.128.
constexpr int REC32 = 12;

template <int KERNEL>
__device__ __forceinline__ void uj_pair32(UJAcc32& a, float tx, float ty, float tz, const float4* __restrict__ rec,
                                          const float2* __restrict__ tab) {
    const float4 s0 = rec[0];
    const float4 s1 = rec[1];
    float dx = tx - s0.x, dy = ty - s0.y, dz = tz - s0.z;
    float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
    float A, B;
    if (KERNEL == K_SINGULAR) {
        float ri = r2 > 0.f ? rsqrtf(r2) : 0.f;
        float ri2 = ri * ri;
        A = ri2 * ri;
        B = (-3.f * ri2) * A;
    } else if (KERNEL == K_GAUSSIANERF) {
        float t = r2 * s0.w;
        if (__all_sync(0xffffffffu, t >= VPM_GT32_TFAR)) {
            float ri = rsqrtf(r2);
            float ri2 = ri * ri;
            A = ri2 * ri;
            B = (-3.f * ri2) * A;
        } else if (t < VPM_GT32_TFAR) {
            float q = rintf(t * (float)VPM_GT_INVW);
            int i = (int)q;
            float u = fmaf(q, -(float)VPM_GT_W, t);
            const float2* tp = tab + i;
            float2 c = tp[VPM_GT32_DEG * VPM_GT32_NINT];
            float G = c.x, H = c.y;
#pragma unroll
            for (int k = VPM_GT32_DEG - 1; k >= 0; --k) {
                c = tp[k * VPM_GT32_NINT];
                G = fmaf(G, u, c.x);
                H = fmaf(H, u, c.y);
            }
            A = r2 > 0.f ? G * s1.w : 0.f;
            B = H * rec[2].x;
        } else {
            float ri = rsqrtf(r2);
            float ri2 = ri * ri;
            A = ri2 * ri;
            B = (-3.f * ri2) * A;
        }
    } else if (KERNEL == K_WINCKELMANS) {
        float t = r2 * s0.w;
        float w = rsqrtf(t + 1.f);
        float w2 = w * w, w4 = w2 * w2, w5 = w4 * w, w7 = w5 * w2;
        A = r2 > 0.f ? ((t + 2.5f) * w5) * s1.w : 0.f;
        B = (fmaf(-3.f, t, -10.5f) * w7) * rec[2].x;
    } else {
        float t = r2 * s0.w;
        float tt = t > 0.f ? t : 1.f;
        float rs = rsqrtf(tt);
        float s3 = tt * (tt * rs);
        float e = __expf(-s3);
        float rs2 = rs * rs;
        float G = (s3 < 1e-3f ? s3 * (1.f - 0.5f * s3) : 1.f - e) * (rs2 * rs);
        float H = (3.f * (e - G)) * rs2;
        A = r2 > 0.f ? G * s1.w : 0.f;
        B = r2 > 0.f ? H * rec[2].x : 0.f;
    }
    uj_accumulate32(a, dx, dy, dz, s1.x, s1.y, s1.z, A, B);
}

struct __align__(16) PairSmem32 {
    double tile[2][TILE_DOUBLES];          // FP64 tiles as they arrive from HBM (bulk TMA)
    float tile32[TILE_SRC * REC32];        // the current tile converted to FP32, CTA-relative
    uint64_t full[2];
};

constexpr size_t uj_f32_smem_bytes(int kernel) {
    return sizeof(PairSmem32) + (kernel == K_GAUSSIANERF ? sizeof(float) * 2 * (VPM_GT32_DEG + 1) * VPM_GT32_NINT : 0);
}

template <int KERNEL>
__global__ void __launch_bounds__(UJ_BT, 2)
uj_direct_f32_kernel(const double* __restrict__ srec, int ntiles, const double* __restrict__ tx,
                     const double* __restrict__ ty, const double* __restrict__ tz, int64_t nt, double* __restrict__ U,
                     double* __restrict__ J, int64_t ldo, int accumulate, const float* __restrict__ gh_table,
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
    PairSmem32& sm = *reinterpret_cast<PairSmem32*>(smem_raw);
    float2* tab = reinterpret_cast<float2*>(smem_raw + sizeof(PairSmem32));

    const int tid = threadIdx.x;
    const int64_t i0 = (int64_t)blockIdx.x * UJ_BT;
    const int64_t i = i0 + tid;

    if (tid == 0) {
        mbar_init(&sm.full[0], 1);
        mbar_init(&sm.full[1], 1);
        fence_mbar_init();
    }
    if (KERNEL == K_GAUSSIANERF) {
        const float2* g2 = reinterpret_cast<const float2*>(gh_table);
        for (int k = tid; k < (VPM_GT32_DEG + 1) * VPM_GT32_NINT; k += UJ_BT) tab[k] = g2[k];
    }
    __syncthreads();
    if (tid == 0 && ntiles > 0) {
        mbar_arrive_expect_tx(&sm.full[0], TILE_BYTES);
        bulk_g2s(sm.tile[0], srec, TILE_BYTES, &sm.full[0]);
    }

    // CTA origin: the first target of the block (always valid since the grid is ceil(nt / UJ_BT))
    const double ox = tx[i0], oy = ty[i0], oz = tz[i0];
    const bool live = i < nt;
    const float px = live ? (float)(tx[i] - ox) : 0.f, py = live ? (float)(ty[i] - oy) : 0.f,
                pz = live ? (float)(tz[i] - oz) : 0.f;
    double tot[15];
#pragma unroll
    for (int c = 0; c < 15; ++c) tot[c] = 0.0;

    for (int k = 0; k < ntiles; ++k) {
        const int b = k & 1;
        if (tid == 0 && k + 1 < ntiles) {
            mbar_arrive_expect_tx(&sm.full[b ^ 1], TILE_BYTES);
            bulk_g2s(sm.tile[b ^ 1], srec + (size_t)(k + 1) * TILE_DOUBLES, TILE_BYTES, &sm.full[b ^ 1]);
        }
        mbar_wait(&sm.full[b], (k >> 1) & 1);
        // convert this tile to FP32 (one record per thread: TILE_SRC == UJ_BT)
        {
            const double* r = sm.tile[b] + tid * REC_REALS;
            float4* o = reinterpret_cast<float4*>(sm.tile32 + tid * REC32);
            // FP64 record { x, y | z, G'x | G'y, G'z | T_FAR s^2, 1/s^3 | 1/s^5, 1/s^2 } -> FP32 working record
            o[0] = make_float4((float)(r[0] - ox), (float)(r[1] - oy), (float)(r[2] - oz), (float)r[9]);
            o[1] = make_float4((float)r[3], (float)r[4], (float)r[5], (float)r[7]);
            o[2] = make_float4((float)r[8], 0.f, 0.f, 0.f);
        }
        __syncthreads();
        const float4* rec = reinterpret_cast<const float4*>(sm.tile32);
        UJAcc32 a = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll 4
        for (int j = 0; j < TILE_SRC; ++j) uj_pair32<KERNEL>(a, px, py, pz, rec + j * (REC32 / 4), tab);
        tot[0] += a.u0; tot[1] += a.u1; tot[2] += a.u2;
        tot[3] += a.j0; tot[4] += a.j1; tot[5] += a.j2; tot[6] += a.j3; tot[7] += a.j4;
        tot[8] += a.j5; tot[9] += a.j6; tot[10] += a.j7; tot[11] += a.j8;
        tot[12] += a.w0; tot[13] += a.w1; tot[14] += a.w2;
        __syncthreads();
    }

    if (live) {
        tot[4] -= tot[14]; tot[5] += tot[13];
        tot[6] += tot[14]; tot[8] -= tot[12];
        tot[9] -= tot[13]; tot[10] += tot[12];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            double* p = U + (size_t)c * ldo + i;
            *p = accumulate ? *p + tot[c] : tot[c];
        }
#pragma unroll
        for (int c = 0; c < 9; ++c) {
            double* p = J + (size_t)c * ldo + i;
            *p = accumulate ? *p + tot[3 + c] : tot[3 + c];
        }
    }
}

static_assert(TILE_SRC == UJ_BT, "the FP32 tile conversion assigns one record per thread");

}  // namespace vpm
