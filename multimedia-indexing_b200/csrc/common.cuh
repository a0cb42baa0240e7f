// common.cuh -- shared device helpers for libmmidx (sm_100a only).
//
// Numeric model: the reference is Java binary64 with one rounding per sub / mul / add and
// index-ascending accumulation starting from 0.0.  Every distance below is built from sqacc(), which uses
// the explicit round-to-nearest intrinsics so nvcc can never contract a*b+c into an FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define MMIDX_NT 256  // threads per CTA for every kernel in this library

namespace mmidx {

// acc + (x - y) * (x - y) with three roundings, exactly as
//   distance += (a[j] - b[j]) * (a[j] - b[j]);       IVFPQ.java:532-533,553,583,619  PQ.java:393,417  Linear.java:148
// (x - y) and (y - x) are exact negations, so operand order inside the square is irrelevant bit-wise.
__device__ __forceinline__ double sqacc(double acc, double x, double y) {
    double a = __dsub_rn(x, y);
    return __dadd_rn(acc, __dmul_rn(a, a));
}

// fp32-safe magnitude window of the fp32 filters (fast_scan.cuh, coarse_fast.cuh): a norm entering a filter is 0 or
// inside [FAST_MAG_MIN, FAST_MAG_MAX]; FAST_ABS_SLACK is the absolute term the error radii carry for underflow.
constexpr double FAST_MAG_MIN = 1e-12, FAST_MAG_MAX = 1e12, FAST_ABS_SLACK = 1e-30;
__host__ __device__ inline bool fast_mag_ok(double v) { return v == 0.0 || (v >= FAST_MAG_MIN && v <= FAST_MAG_MAX); }

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) -----------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// bounded spin: a mis-programmed barrier traps instead of hanging the GPU box
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 26)) __trap();
    }
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes % 16 == 0; completes on `bar`
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// streaming 128-bit load (codes are read once per scan; keep them out of L1)
__device__ __forceinline__ uint4 ld_nc_u4(const void *p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// volatile 128-bit load: ptxas keeps volatile accesses in program order, which pins this load ABOVE the (volatile)
// shared-memory lookups of the current step -- a plain ld.global.nc gets sunk below them to save registers
__device__ __forceinline__ uint4 ld_vol_u4(const void *p) {
    uint4 r;
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ uint2 ld_nc_u2(const void *p) {
    uint2 r;
    asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
    return r;
}

}  // namespace mmidx
