// Inline-PTX wrappers for the sm_100a tensor path (tcgen05 / TMEM / TMA / mbarrier) shared by the
// tensor-engine kernels.  Everything here is a thin, documented spelling of one PTX instruction.
#pragma once

#include <cuda.h>

#include "common.cuh"

namespace chaorec {
namespace ptx {

__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, int c0, int c1, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// try_wait with a suspend-time hint: the thread sleeps in hardware until the phase completes (or the hint
// expires) instead of spinning through the issue slots of the epilogue warps on the same scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint64_t *bar, uint32_t parity) {
    uint32_t ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(ok)
            : "r"(smem_u32(bar)), "r"(parity), "r"(100000u)
            : "memory");
    }
}
// true for exactly one lane of a converged warp (SASS: ELECT).  Code guarded by it keeps its operands in
// uniform registers; `if (lane == 0)` around a tcgen05.mma makes the compiler emit a per-instruction
// R2UR / BRA.U.ANY "waterfall" loop instead (measured: 72 clk per MMA against 34, profiles/r02_probe_*).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T : the "TS" form.  A sits in tensor memory with row = lane and two 16-bit
// elements per 32-bit column (the layout tcgen05.st.32x32b produces when every thread writes its own row).
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
#define CHAOREC_R32_LIST                                                                                             \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
    "%24, %25, %26, %27, %28, %29, %30, %31}"
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " CHAOREC_R32_LIST ", [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
                   "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
                   "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], " CHAOREC_R32_LIST ";" ::"r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
                 "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]),
                 "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]), "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]),
                 "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]), "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31]),
                 "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};" ::"r"(v[0]),
        "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]), "r"(v[10]), "r"(v[11]),
        "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// shared-memory matrix descriptor, K-major operand rows of 128 bytes with the 128-byte swizzle:
// stride between 8-row groups 1024 B, descriptor version 1 (sm_100), SWIZZLE_128B
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);
__device__ __forceinline__ uint64_t smem_desc_sw128(const void *p) {
    return (static_cast<uint64_t>(kDescHiSw128) << 32) | static_cast<uint64_t>((smem_u32(p) >> 4) & 0x3FFFu);
}
// tcgen05 instruction descriptor (kind::f16): c_fmt 1 = F32 / 0 = F16 accumulators; ab_fmt 1 = BF16 / 0 = F16
// operands; K-major A and B; M = 128.
__host__ __device__ constexpr uint32_t make_idesc_m128(int n, int c_fmt, int ab_fmt) {
    return (static_cast<uint32_t>(c_fmt) << 4) | (static_cast<uint32_t>(ab_fmt) << 7) | (static_cast<uint32_t>(ab_fmt) << 10) |
           ((static_cast<uint32_t>(n) >> 3) << 17) | ((128u >> 4) << 24);
}

// warp-wide bitonic sort of P (power of two) 64-bit keys in shared memory, descending
__device__ __forceinline__ void warp_sort_desc(uint64_t *s, int P, int lane) {
    for (int k = 2; k <= P; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = lane; t < (P >> 1); t += 32) {
                const int i = ((t & ~(j - 1)) << 1) | (t & (j - 1));
                const int l = i | j;
                const bool desc = ((i & k) == 0);
                const uint64_t a = s[i], b = s[l];
                if ((a < b) == desc && a != b) {
                    s[i] = b;
                    s[l] = a;
                }
            }
            __syncwarp();
        }
    }
}

// The same network for exactly 64 keys held two per lane (element e = lane + 32 r): compare-exchanges go through
// shuffles instead of shared memory -- ~300 warp instructions against ~730 for the shared-memory form.
__device__ __forceinline__ void warp_sort64_desc_regs(uint64_t &k0, uint64_t &k1, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        // direction of the length-k run element e sits in: bit k of e; for k = 32 that bit is r, not a lane bit
        const bool desc0 = (k == 32) ? true : ((lane & k) == 0);
        const bool desc1 = (k == 32) ? false : desc0;
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool lower = (lane & j) == 0;
            const uint64_t o0 = __shfl_xor_sync(0xffffffffu, k0, j);
            const uint64_t o1 = __shfl_xor_sync(0xffffffffu, k1, j);
            k0 = (lower == desc0) ? (k0 > o0 ? k0 : o0) : (k0 < o0 ? k0 : o0);
            k1 = (lower == desc1) ? (k1 > o1 ? k1 : o1) : (k1 < o1 ? k1 : o1);
        }
    }
    // k = 64: descending everywhere; j = 32 pairs the lane's own two elements
    {
        const uint64_t hi = k0 > k1 ? k0 : k1, lo = k0 > k1 ? k1 : k0;
        k0 = hi;
        k1 = lo;
    }
#pragma unroll
    for (int j = 16; j > 0; j >>= 1) {
        const bool keep_max = (lane & j) == 0;
        const uint64_t o0 = __shfl_xor_sync(0xffffffffu, k0, j);
        const uint64_t o1 = __shfl_xor_sync(0xffffffffu, k1, j);
        k0 = keep_max ? (k0 > o0 ? k0 : o0) : (k0 < o0 ? k0 : o0);
        k1 = keep_max ? (k1 > o1 ? k1 : o1) : (k1 < o1 ? k1 : o1);
    }
}
// 32 keys, one per lane
__device__ __forceinline__ void warp_sort32_desc_regs(uint64_t &k0, int lane) {
#pragma unroll
    for (int k = 2; k <= 32; k <<= 1) {
        const bool desc = (k == 32) ? true : ((lane & k) == 0);
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            const bool lower = (lane & j) == 0;
            const uint64_t o0 = __shfl_xor_sync(0xffffffffu, k0, j);
            k0 = (lower == desc) ? (k0 > o0 ? k0 : o0) : (k0 < o0 ? k0 : o0);
        }
    }
}
// sort P keys in shared memory (descending): through registers when P == 64
__device__ __forceinline__ void warp_sort_desc_fast(uint64_t *s, int P, int lane) {
    if (P == 64) {
        uint64_t k0 = s[lane], k1 = s[lane + 32];
        warp_sort64_desc_regs(k0, k1, lane);
        s[lane] = k0;
        s[lane + 32] = k1;
        __syncwarp();
    } else {
        warp_sort_desc(s, P, lane);
    }
}

}  // namespace ptx
}  // namespace chaorec
