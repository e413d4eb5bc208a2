// uj_direct_f32.cuh — FP32 variant of K1 (the reference's knob is `vpm_floattype`,
// /root/reference/src/FLOWUnsteady_simulation.jl:137).  Pair arithmetic and per-tile partial sums are FP32; the state,
// the source tiles in HBM and the cross-tile totals stay FP64.
//
// Blackwell-specific: the FP32 path is issue-bound, not pipe-bound (~60 issue slots per pair at one FFMA each), so every
// thread owns TWO targets and the pair arithmetic runs on packed `fma.rn.f32x2` (SASS FFMA2: two FMAs per issue slot,
// sm_100+).  Tiles arrive as FP64 records by bulk TMA (same wire format as the FP64 kernel), and are converted once per
// tile into an FP32 working tile whose fields are already duplicated / negated the way the packed arithmetic consumes
// them ({-x,-x,-y,-y}, {-z,-z,gx,gx}, ...), relative to the CTA's first target so close pairs keep their separation digits.
// Far tiles (tile-level box test, as in the FP64 kernel) run a branch-free packed loop; near / mixed tiles compute A, B
// per target half in scalar arithmetic with the FP32 G/H table and share the packed accumulation.
#pragma once

#include "uj_direct.cuh"

namespace vpm {

struct UJAcc32x2 {   // .x = first target of the thread, .y = second
    float2 u0, u1, u2, j0, j1, j2, j3, j4, j5, j6, j7, j8, w0, w1, w2;
};

__device__ __forceinline__ void acc32x2_zero(UJAcc32x2& a) {
    const float2 z = make_float2(0.f, 0.f);
    a.u0 = a.u1 = a.u2 = z;
    a.j0 = a.j1 = a.j2 = a.j3 = a.j4 = a.j5 = a.j6 = a.j7 = a.j8 = z;
    a.w0 = a.w1 = a.w2 = z;
}

// FP32 working record: 6 float4 = 24 floats per source
//   q0 {-x,-x,-y,-y}  q1 {-z,-z, gx, gx}  q2 { gy, gy, gz, gz}  q3 {-gx,-gx,-gy,-gy}  q4 {-gz,-gz, T_far s^2, 1/s^2}  q5 {1/s^3, 1/s^5, 0, 0}
constexpr int REC32 = 24;
constexpr int F32_TPB = 2 * UJ_BT;   // targets per CTA

// tail of one interaction for BOTH targets once A = g/r^3 and B = (g'/(sigma r) - 3 g/r^2)/r^3 are known (packed)
__device__ __forceinline__ void uj_accumulate32x2(UJAcc32x2& a, float2 dx, float2 dy, float2 dz, float2 gx, float2 gy,
                                                  float2 gz, float2 ngx, float2 ngy, float2 ngz, float2 A, float2 B) {
    float2 c0 = __ffma2_rn(dy, gz, __fmul2_rn(dz, ngy));
    float2 c1 = __ffma2_rn(dz, gx, __fmul2_rn(dx, ngz));
    float2 c2 = __ffma2_rn(dx, gy, __fmul2_rn(dy, ngx));
    a.u0 = __ffma2_rn(A, c0, a.u0);
    a.u1 = __ffma2_rn(A, c1, a.u1);
    a.u2 = __ffma2_rn(A, c2, a.u2);
    float2 b0 = __fmul2_rn(B, c0), b1 = __fmul2_rn(B, c1), b2 = __fmul2_rn(B, c2);
    a.j0 = __ffma2_rn(b0, dx, a.j0); a.j1 = __ffma2_rn(b1, dx, a.j1); a.j2 = __ffma2_rn(b2, dx, a.j2);
    a.j3 = __ffma2_rn(b0, dy, a.j3); a.j4 = __ffma2_rn(b1, dy, a.j4); a.j5 = __ffma2_rn(b2, dy, a.j5);
    a.j6 = __ffma2_rn(b0, dz, a.j6); a.j7 = __ffma2_rn(b1, dz, a.j7); a.j8 = __ffma2_rn(b2, dz, a.j8);
    a.w0 = __ffma2_rn(A, gx, a.w0);
    a.w1 = __ffma2_rn(A, gy, a.w1);
    a.w2 = __ffma2_rn(A, gz, a.w2);
}

// A, B of one target half (scalar; used in near / mixed tiles and for the non-Gaussian kernels)
template <int KERNEL>
__device__ __forceinline__ void ab32(float r2, float rfar2, float sinv2, float sinv3, float sinv5,
                                     const float2* __restrict__ tab, float& A, float& B) {
    if (KERNEL == K_SINGULAR || (KERNEL == K_GAUSSIANERF && r2 > rfar2)) {
        float ri = r2 > 0.f ? rsqrtf(r2) : 0.f;
        float ri2 = ri * ri;
        A = ri2 * ri;
        B = (-3.f * ri2) * A;
    } else if (KERNEL == K_GAUSSIANERF) {
        float t = r2 * sinv2;
        float q = rintf(t * (float)VPM_GT_INVW);
        int i = min((int)q, VPM_GT32_NINT - 1);
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
        A = r2 > 0.f ? G * sinv3 : 0.f;
        B = H * sinv5;
    } else if (KERNEL == K_WINCKELMANS) {
        float t = r2 * sinv2;
        float w = rsqrtf(t + 1.f);
        float w2 = w * w, w4 = w2 * w2, w5 = w4 * w, w7 = w5 * w2;
        A = r2 > 0.f ? ((t + 2.5f) * w5) * sinv3 : 0.f;
        B = (fmaf(-3.f, t, -10.5f) * w7) * sinv5;
    } else {   // K_GAUSSIAN
        float t = r2 * sinv2;
        float tt = t > 0.f ? t : 1.f;
        float rs = rsqrtf(tt);
        float s3 = tt * (tt * rs);
        float e = __expf(-s3);
        float rs2 = rs * rs;
        float G = (s3 < 1e-3f ? s3 * (1.f - 0.5f * s3) : 1.f - e) * (rs2 * rs);
        float H = (3.f * (e - G)) * rs2;
        A = r2 > 0.f ? G * sinv3 : 0.f;
        B = r2 > 0.f ? H * sinv5 : 0.f;
    }
}

struct __align__(16) PairSmem32 {
    double tile[2][TILE_DOUBLES];          // FP64 tiles as they arrive from HBM (bulk TMA)
    float tile32[TILE_SRC * REC32];        // the current tile converted to FP32, CTA-relative, duplicated for FFMA2
    uint64_t full[2];
    double tbox[8];                        // bounding box of the CTA's targets
    double red[6][8];
};

constexpr size_t uj_f32_smem_bytes(int kernel) {
    return sizeof(PairSmem32) + (kernel == K_GAUSSIANERF ? sizeof(float) * 2 * (VPM_GT32_DEG + 1) * VPM_GT32_NINT : 0);
}

// Grid: (ceil(nt / F32_TPB), nchunks).  Thread t owns targets i0 + t and i0 + UJ_BT + t.
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
    const int64_t i0 = (int64_t)blockIdx.x * F32_TPB;
    const int64_t ia = i0 + tid, ib = i0 + UJ_BT + tid;

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

    // CTA origin: the first target of the block (always valid since the grid is ceil(nt / F32_TPB))
    const double ox = tx[i0], oy = ty[i0], oz = tz[i0];
    const bool la = ia < nt, lb = ib < nt;
    const double ax = la ? tx[ia] : ox, ay = la ? ty[ia] : oy, az = la ? tz[ia] : oz;
    const double bx = lb ? tx[ib] : ox, by = lb ? ty[ib] : oy, bz = lb ? tz[ib] : oz;
    const float2 px = make_float2((float)(ax - ox), (float)(bx - ox));
    const float2 py = make_float2((float)(ay - oy), (float)(by - oy));
    const float2 pz = make_float2((float)(az - oz), (float)(bz - oz));
    if (KERNEL == K_GAUSSIANERF) {   // bounding box of the CTA's targets (absolute FP64 coordinates)
        const double big = 1.0e300;
        double v[6] = {fmin(la ? ax : big, lb ? bx : big), fmin(la ? -ax : big, lb ? -bx : big),
                       fmin(la ? ay : big, lb ? by : big), fmin(la ? -ay : big, lb ? -by : big),
                       fmin(la ? az : big, lb ? bz : big), fmin(la ? -az : big, lb ? -bz : big)};
#pragma unroll
        for (int c = 0; c < 6; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v[c] = fmin(v[c], __shfl_xor_sync(0xffffffffu, v[c], o));
        if ((tid & 31) == 0)
            for (int c = 0; c < 6; ++c) sm.red[c][tid >> 5] = v[c];
        __syncthreads();
        if (tid < 6) {
            double m = sm.red[tid][0];
            for (int k = 1; k < UJ_BT / 32; ++k) m = fmin(m, sm.red[tid][k]);
            sm.tbox[tid] = (tid & 1) ? -m : m;
        }
        __syncthreads();
    }
    double tot[30];
#pragma unroll
    for (int c = 0; c < 30; ++c) tot[c] = 0.0;

    for (int k = 0; k < ntiles; ++k) {
        const int b = k & 1;
        if (tid == 0 && k + 1 < ntiles) {
            mbar_arrive_expect_tx(&sm.full[b ^ 1], TILE_BYTES);
            bulk_g2s(sm.tile[b ^ 1], srec + (size_t)(k + 1) * TILE_DOUBLES, TILE_BYTES, &sm.full[b ^ 1]);
        }
        mbar_wait(&sm.full[b], (k >> 1) & 1);
        {   // convert this tile to the FP32 working layout (one record per thread: TILE_SRC == UJ_BT)
            // FP64 record { x, y | z, G'x | G'y, G'z | T_far s^2, 1/s^3 | 1/s^5, 1/s^2 }
            const double* r = sm.tile[b] + tid * REC_REALS;
            const float x = (float)(r[0] - ox), y = (float)(r[1] - oy), z = (float)(r[2] - oz);
            const float gx = (float)r[3], gy = (float)r[4], gz = (float)r[5];
            float4* o = reinterpret_cast<float4*>(sm.tile32 + tid * REC32);
            o[0] = make_float4(-x, -x, -y, -y);
            o[1] = make_float4(-z, -z, gx, gx);
            o[2] = make_float4(gy, gy, gz, gz);
            o[3] = make_float4(-gx, -gx, -gy, -gy);
            o[4] = make_float4(-gz, -gz, (float)r[6], (float)r[9]);
            o[5] = make_float4((float)r[7], (float)r[8], 0.f, 0.f);
        }
        bool tile_far = false;
        if (KERNEL == K_GAUSSIANERF) {
            // FP32 reaches |1 - g| < 3e-8 already at t = 40; the FP64 kernel's (larger) radius is used: conservative
            const double* hdr = sm.tile[b] + TILE_HDR;
            tile_far = box_dist2(sm.tbox, hdr) > hdr[6];
        }
        __syncthreads();
        const float4* rec = reinterpret_cast<const float4*>(sm.tile32);
        UJAcc32x2 a;
        acc32x2_zero(a);
        if (tile_far) {
            // every pair of this tile is in the singular regime (r2 > 0 guaranteed): branch-free packed loop (FFMA2)
#pragma unroll 4
            for (int j = 0; j < TILE_SRC; ++j) {
                const float4 q0 = rec[j * 6 + 0], q1 = rec[j * 6 + 1], q2 = rec[j * 6 + 2], q3 = rec[j * 6 + 3], q4 = rec[j * 6 + 4];
                const float2 dx = __fadd2_rn(px, make_float2(q0.x, q0.y));
                const float2 dy = __fadd2_rn(py, make_float2(q0.z, q0.w));
                const float2 dz = __fadd2_rn(pz, make_float2(q1.x, q1.y));
                const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                const float2 ri = make_float2(rsqrtf(r2.x), rsqrtf(r2.y));
                const float2 ri2 = __fmul2_rn(ri, ri);
                const float2 A = __fmul2_rn(ri2, ri);
                const float2 B = __fmul2_rn(__fmul2_rn(ri2, make_float2(-3.f, -3.f)), A);
                uj_accumulate32x2(a, dx, dy, dz, make_float2(q1.z, q1.w), make_float2(q2.x, q2.y), make_float2(q2.z, q2.w),
                                  make_float2(q3.x, q3.y), make_float2(q3.z, q3.w), make_float2(q4.x, q4.y), A, B);
            }
        } else {
#pragma unroll 2
            for (int j = 0; j < TILE_SRC; ++j) {
                const float4 q0 = rec[j * 6 + 0], q1 = rec[j * 6 + 1], q2 = rec[j * 6 + 2], q3 = rec[j * 6 + 3], q4 = rec[j * 6 + 4],
                             q5 = rec[j * 6 + 5];
                const float2 dx = __fadd2_rn(px, make_float2(q0.x, q0.y));
                const float2 dy = __fadd2_rn(py, make_float2(q0.z, q0.w));
                const float2 dz = __fadd2_rn(pz, make_float2(q1.x, q1.y));
                const float2 r2 = __ffma2_rn(dz, dz, __ffma2_rn(dy, dy, __fmul2_rn(dx, dx)));
                float2 A, B;
                // FP32 far radius t >= 40 (|1 - g| < 3e-8); q4.z carries the FP64 radius T_far sigma^2 = 88 sigma^2
                const float rfar2 = q4.z * (VPM_GT32_TFAR / (float)VPM_GT_TFAR);
                ab32<KERNEL>(r2.x, rfar2, q4.w, q5.x, q5.y, tab, A.x, B.x);
                ab32<KERNEL>(r2.y, rfar2, q4.w, q5.x, q5.y, tab, A.y, B.y);
                uj_accumulate32x2(a, dx, dy, dz, make_float2(q1.z, q1.w), make_float2(q2.x, q2.y), make_float2(q2.z, q2.w),
                                  make_float2(q3.x, q3.y), make_float2(q3.z, q3.w), make_float2(q4.x, q4.y), A, B);
            }
        }
        const float2* av = reinterpret_cast<const float2*>(&a);
#pragma unroll
        for (int c = 0; c < 15; ++c) {
            tot[c] += av[c].x;
            tot[15 + c] += av[c].y;
        }
        __syncthreads();
    }

#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const bool live = half ? lb : la;
        const int64_t i = half ? ib : ia;
        double* t = tot + 15 * half;   // u0..u2, j0..j8, w0..w2
        if (live) {
            t[4] -= t[14]; t[5] += t[13];
            t[6] += t[14]; t[8] -= t[12];
            t[9] -= t[13]; t[10] += t[12];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                double* p = U + (size_t)c * ldo + i;
                *p = accumulate ? *p + t[c] : t[c];
            }
#pragma unroll
            for (int c = 0; c < 9; ++c) {
                double* p = J + (size_t)c * ldo + i;
                *p = accumulate ? *p + t[3 + c] : t[3 + c];
            }
        }
    }
}

static_assert(TILE_SRC == UJ_BT, "the FP32 tile conversion assigns one record per thread");

}  // namespace vpm
