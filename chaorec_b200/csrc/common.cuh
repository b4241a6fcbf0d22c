// Shared device/host helpers for the sm_100a kernels of chaorec_b200.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/chaorec_b200.h"

#define CHAOREC_NUM_SMS_FALLBACK 148

namespace chaorec {

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? CHAOREC_OK : CHAOREC_E_LAUNCH;
}

inline int sm_count() {
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else {
            cudaGetLastError();
            cached = CHAOREC_NUM_SMS_FALLBACK;
        }
    }
    return cached;
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// ---- orderable keys -----------------------------------------------------------------------
// key = (orderable(score) << 32) | (0xFFFFFFFF - index): a larger key is a better candidate
// under "descending score, ties -> ascending index".  -0.0f is folded into +0.0f first so that
// the order agrees with a float comparison (torch.sort treats them as equal).
__host__ __device__ __forceinline__ uint32_t orderable_bits(float s) {
    s = s + 0.0f;
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(s);
#else
    union { float f; uint32_t u; } cv; cv.f = s; uint32_t b = cv.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float score_of_bits(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float s, uint32_t idx) {
    return (static_cast<uint64_t>(orderable_bits(s)) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}
__host__ __device__ __forceinline__ uint32_t key_index(uint64_t key) {
    return 0xFFFFFFFFu - static_cast<uint32_t>(key & 0xFFFFFFFFull);
}
__host__ __device__ __forceinline__ float key_score(uint64_t key) {
    return score_of_bits(static_cast<uint32_t>(key >> 32));
}

#ifdef __CUDACC__
// ---- mbarrier + bulk async copy (TMA, descriptor-less 1-D form) ----------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
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
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
// global -> shared bulk copy completing on an mbarrier (SASS: UBLKCP).  dst/src 16-byte
// aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
#endif  // __CUDACC__

}  // namespace chaorec
