// common.cuh — shared definitions for the vpmb200 engine (sm_100a only).
//
// Device-resident particle state is SoA: field f of particle i lives at state[f * ld + i], with the
// same 43 fields (and order) as the reference's 43 x np particle matrix (matrix
// addressing confirmed at /root/reference/src/FLOWUnsteady_simulation.jl:509-510).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace vpm {

// ---- particle record (0-based field offsets; identical to oracle/vpm_oracle.h) -------------------
constexpr int NFIELDS = 43;
constexpr int F_X = 0, F_GAMMA = 3, F_SIGMA = 6, F_VOL = 7, F_CIRC = 8, F_U = 9, F_W = 12, F_J = 15,
              F_PSE = 24, F_M = 27, F_C = 36, F_SFS = 39, F_STATIC = 42;

enum KernelId { K_GAUSSIANERF = 0, K_WINCKELMANS = 1, K_GAUSSIAN = 2, K_SINGULAR = 3 };

// 1/(4 pi) exactly as the reference writes it (src/FLOWUnsteady_processing_force.jl:903)
constexpr double CONST4 = 0.07957747154594767;
// 1/(2 pi)^(3/2) : zeta_gaussianerf(0)
constexpr double CONST1 = 0.06349363593424097;

// ---- source records streamed through shared memory by the pairwise kernels ------------------------
// 10 doubles (80 B) per source, 16-byte aligned so they move as LDS.128 broadcasts:
//   UJ pass  : { x, y | z, G'x | G'y, G'z | T_FAR sigma^2, 1/sigma^3 | 1/sigma^5, 1/sigma^2 },  G' = -Gamma/(4 pi)
//              (the far-field branch reads only the first three quads)
//   E_str    : { x, y | z, 1/sigma^2 | cGx, cGy | cGz, c vx | c vy, c vz },  c = zeta(0)/sigma^3 (kernel
//              normalisation folded in), v = J^T Gamma (transposed scheme) or J Gamma (classic)
// A tile is TILE_SRC records followed by one 10-double header { xmin, xmax, ymin, ymax, zmin, zmax,
// max(T_FAR sigma^2), n_real, 0, 0 } (bounding box of the tile's real sources), so a tile moves with ONE bulk copy
// and tiles from different ranks concatenate (the multi-GPU direct path all-gathers whole tiles).
// Tail slots of the last tile repeat the position of the tile's first real source with zero strength: they
// contribute exactly 0 on every branch and never sit closer to a target than a real source does.
constexpr int REC_REALS = 10;
constexpr int TILE_SRC = 256;                              // sources per shared-memory tile
constexpr int TILE_DOUBLES = (TILE_SRC + 1) * REC_REALS;   // 2570 doubles = 20,560 B per tile incl. header
constexpr int TILE_HDR = TILE_SRC * REC_REALS;             // offset of the header inside a tile

// ---- PTX helpers: mbarrier + 1-D bulk TMA copy (cp.async.bulk, SASS UBLKCP) -----------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ---- FP64 reciprocal square root: MUFU.RSQ64H seed (2^-22) + one third-order step (5 FP64 ops) -----
__device__ __forceinline__ double rsqrt_f64(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double h = x * y;
    double e = fma(-h, y, 1.0);
    double p = fma(0.375, e, 0.5);
    return fma(y * e, p, y);
}

}  // namespace vpm
