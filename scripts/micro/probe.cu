// Micro-probes that settle design questions for the evaluation kernel on a real B200 (run through gpurun):
//   (1) issue rates of the candidate filter instructions (FMNMX3, VHMNMX, VIMNMX3.S16x2, HSETP2, LOP3, FFMA)
//   (2) tcgen05.mma dispatch cost: A from shared memory (SS) vs A from tensor memory (TS), N = 64/128/256,
//       fp32 vs fp16 accumulators
//   (3) correctness of the TS form with A written by tcgen05.st.32x32b (row = lane, two 16-bit elements per
//       column) and of fp16 accumulators read back with tcgen05.ld ... .pack::16b, incl. the rounding mode
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/micro/probe.bin scripts/micro/probe.cu
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#define CK(x)                                                                                   \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
#define LD32_REGS(v)                                                                                                       \
    "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), \
        "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]),            \
        "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),            \
        "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
#define ST32_REGS(v)                                                                                                     \
    "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),         \
        "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),      \
        "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),      \
        "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
#define R32_LIST                                                                                                     \
    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, " \
    "%24, %25, %26, %27, %28, %29, %30, %31}"
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 " R32_LIST ", [%32];" : LD32_REGS(v) : "r"(taddr) : "memory");
}
// 64 columns of 16-bit accumulators (one per 32-bit column) packed two per register
__device__ __forceinline__ void tmem_ld64_pack16(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.pack::16b.b32 " R32_LIST ", [%32];" : LD32_REGS(v) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "
        "%24, %25, %26, %27, %28, %29, %30, %31};" ::ST32_REGS(v),
        "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO = 1024 B, version 1, SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc(const void *p) {
    return (static_cast<uint64_t>(kDescHi) << 32) | static_cast<uint64_t>((smem_u32(p) >> 4) & 0x3FFFu);
}
// c_fmt: 1 = F32, 0 = F16;  ab_fmt: 1 = BF16, 0 = F16
__host__ __device__ constexpr uint32_t make_idesc(int n, int c_fmt, int ab_fmt) {
    return (static_cast<uint32_t>(c_fmt) << 4) | (static_cast<uint32_t>(ab_fmt) << 7) | (static_cast<uint32_t>(ab_fmt) << 10) |
           ((static_cast<uint32_t>(n) >> 3) << 17) | ((128u >> 4) << 24);
}

// ------------------------------------------------------------------------------------------------------
// (1) instruction issue rates
// ------------------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(512) alu_probe(const uint32_t *in, uint32_t *out, int iters) {
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = in[threadIdx.x * 8 + i];
    const uint32_t x = in[4096 + (threadIdx.x & 7)], y = in[4100 + (threadIdx.x & 3)];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t p = a[(i + 1) & 7], q = a[(i + 3) & 7];
                if (OP == 0) {   // FMNMX3
                    a[i] = __float_as_uint(fmaxf(fmaxf(__uint_as_float(a[i]), __uint_as_float(p)), __uint_as_float(q)));
                } else if (OP == 1) {   // VHMNMX
                    __half2 h = __hmax2(__hmax2(*reinterpret_cast<__half2 *>(&a[i]), *reinterpret_cast<const __half2 *>(&p)),
                                        *reinterpret_cast<const __half2 *>(&q));
                    a[i] = *reinterpret_cast<uint32_t *>(&h);
                } else if (OP == 2) {   // VIMNMX3.S16x2
                    a[i] = __vimax3_s16x2(a[i], p, q);
                } else if (OP == 3) {   // LOP3
                    a[i] = (a[i] & p) ^ q;
                } else if (OP == 4) {   // FFMA
                    a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), __uint_as_float(x), __uint_as_float(y)));
                } else if (OP == 5) {   // HSET2 / HSETP2-style packed compare folded into a counter
                    const __half2 h = *reinterpret_cast<__half2 *>(&a[i]);
                    const __half2 t = *reinterpret_cast<const __half2 *>(&p);
                    a[i] += __hbgt2(h, t) ? x : y;
                } else if (OP == 6) {   // FMNMX3 and FFMA interleaved (do the two pipes overlap?)
                    if (i & 1)
                        a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), __uint_as_float(x), __uint_as_float(y)));
                    else
                        a[i] = __float_as_uint(fmaxf(fmaxf(__uint_as_float(a[i]), __uint_as_float(p)), __uint_as_float(q)));
                } else if (OP == 7) {   // FSETP + predicated add (one compare per score)
                    if (__uint_as_float(a[i]) > __uint_as_float(p)) a[i] += x;
                } else if (OP == 8) {   // VHMNMX and FFMA interleaved
                    if (i & 1)
                        a[i] = __float_as_uint(fmaf(__uint_as_float(a[i]), __uint_as_float(x), __uint_as_float(y)));
                    else {
                        __half2 h = __hmax2(__hmax2(*reinterpret_cast<__half2 *>(&a[i]), *reinterpret_cast<const __half2 *>(&p)),
                                            *reinterpret_cast<const __half2 *>(&q));
                        a[i] = *reinterpret_cast<uint32_t *>(&h);
                    }
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run_alu(const char *name, const uint32_t *in, uint32_t *out, int sms, double ops_per_inner) {
    const int iters = 4096;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    alu_probe<OP><<<sms, 512>>>(in, out, 64);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    alu_probe<OP><<<sms, 512>>>(in, out, iters);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    // warp instructions (of the probed kind) per SM: 16 warps x iters x 64
    const double winstr = 16.0 * iters * 64.0 * ops_per_inner;
    printf("alu %-28s %8.3f ms  -> %6.3f warp-instr/ns/SM (x1.965 GHz => %5.2f per clk per SM, %5.3f per SMSP)\n", name, ms,
           winstr / (ms * 1e6), winstr / (ms * 1e6) / 1.965, winstr / (ms * 1e6) / 1.965 / 4);
}

// ------------------------------------------------------------------------------------------------------
// (2) tcgen05.mma dispatch cost
// ------------------------------------------------------------------------------------------------------
// One CTA per SM, thread 0 issues: per "tile" RB row blocks x 4 K-steps of a 128 x N x 16 MMA (the evaluation
// kernel's inner loop at d = 64), accumulators rotate over the TMEM columns; one commit per tile, the issuing
// thread stays two tiles ahead of completion.
// ISSUE: 0 = `if (threadIdx.x == 0)` around the whole loop (round-1 style), 1 = warp-uniform loop, elected lane issues,
//        2 = like 1 with two issuing warps (row blocks 0-1 / 2-3)
// ORDER: 0 = row block outer, K inner (4 consecutive MMAs chain into one accumulator); 1 = K outer, row block inner
template <bool TS, int N, int CFMT, int ISSUE, int ORDER>
__global__ void __launch_bounds__(128, 1) mma_probe(int tiles, uint32_t seed) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sa = smem;
    unsigned char *sb = smem + 65536;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 65536 + 256 * 128);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 4);
    uint32_t s = seed + threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    for (int i = threadIdx.x; i < (65536 + 256 * 128) / 4; i += blockDim.x) {
        s = s * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u | (s & 0x83FF83FFu);
    }
    constexpr int n_issuers = ISSUE == 2 ? 2 : 1;
    if (threadIdx.x == 0) {
        mbar_init(&bars[0], n_issuers);
        mbar_init(&bars[1], n_issuers);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    constexpr int RB = 4;
    constexpr uint32_t idesc = make_idesc(N, CFMT, CFMT == 0 ? 0 : 1);
    constexpr int acc0 = TS ? 128 : 0;
    constexpr int nslots = (512 - acc0) / N;       // accumulator slots of N columns
    const int warp = threadIdx.x >> 5;
    const bool issuer = ISSUE == 0 ? (threadIdx.x == 0) : (warp < n_issuers);
    if (issuer) {
        const int rb_lo = ISSUE == 2 ? warp * 2 : 0, rb_hi = ISSUE == 2 ? warp * 2 + 2 : RB;
        uint32_t ph[2] = {0, 0};
        const uint64_t db = smem_desc(sb);
        for (int t = 0; t < tiles; ++t) {
            if (t >= 2) {
                mbar_wait(&bars[t & 1], ph[t & 1]);
                ph[t & 1] ^= 1u;
            }
            tc_fence_after();
            const bool lead = ISSUE == 0 ? true : elect_one();
            if (lead) {
                if (ORDER == 0) {
#pragma unroll
                    for (int rb = rb_lo; rb < rb_hi; ++rb) {
                        const uint32_t d = tb + acc0 + ((t * RB + rb) % nslots) * N;
                        const uint64_t da = smem_desc(sa + rb * 16384);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (TS) umma_ts(d, tb + rb * 32 + k * 8, db + 2 * k, idesc, k > 0);
                            else umma_ss(d, da + 2 * k, db + 2 * k, idesc, k > 0);
                        }
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
#pragma unroll
                        for (int rb = rb_lo; rb < rb_hi; ++rb) {
                            const uint32_t d = tb + acc0 + ((t * RB + rb) % nslots) * N;
                            const uint64_t da = smem_desc(sa + rb * 16384);
                            if (TS) umma_ts(d, tb + rb * 32 + k * 8, db + 2 * k, idesc, k > 0);
                            else umma_ss(d, da + 2 * k, db + 2 * k, idesc, k > 0);
                        }
                    }
                }
                umma_commit(&bars[t & 1]);
            }
            if (ISSUE != 0) __syncwarp();
        }
        for (int t = tiles; t < tiles + 2; ++t)
            if (t >= 2) {
                mbar_wait(&bars[t & 1], ph[t & 1]);
                ph[t & 1] ^= 1u;
            }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <bool TS, int N, int CFMT, int ISSUE, int ORDER>
void run_mma(const char *name, int sms) {
    const int tiles = 40000;
    const size_t smem = 65536 + 256 * 128 + 128;
    auto kern = mma_probe<TS, N, CFMT, ISSUE, ORDER>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    kern<<<sms, 128, smem>>>(2000, 1u);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<sms, 128, smem>>>(tiles, 7u);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = 16.0 * tiles;
    const double flops = mmas * 2.0 * 128 * N * 16 * sms;
    printf("mma %-44s issue %d order %d %8.3f ms  %7.1f ns/MMA (%6.1f clk @1.965)  %7.1f TFLOP/s\n", name, ISSUE, ORDER, ms,
           ms * 1e6 / mmas, ms * 1e6 / mmas * 1.965, flops / (ms * 1e-3) / 1e12);
    fflush(stdout);
}

// The evaluation kernel's issue pattern: NI issuing warps, warp w owns row block w (A at TMEM columns 32*w) and DEPTH
// private accumulator slots; one commit per tile of 4 MMAs (K = 64); the warp waits for the commit DEPTH tiles back
// before it reuses a slot.  No TMA, no epilogue: the B slab is static shared memory.
template <int NI, int DEPTH, bool ONE_LANE_WAIT>
__global__ void __launch_bounds__(128, 1) mma_probe2(int tiles, uint32_t seed) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sb = smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + 8192);     // [NI][DEPTH]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bars + 16);
    uint32_t s = seed + threadIdx.x * 2654435761u + blockIdx.x * 40503u;
    for (int i = threadIdx.x; i < 8192 / 4; i += blockDim.x) {
        s = s * 1664525u + 1013904223u;
        reinterpret_cast<uint32_t *>(smem)[i] = 0x3C003C00u | (s & 0x83FF83FFu);
    }
    if (threadIdx.x == 0) {
        for (int i = 0; i < NI * DEPTH; ++i) mbar_init(&bars[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) tmem_alloc(tmem_slot, 512);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t idesc = make_idesc(64, 1, 1);
    if (warp < NI) {
        const uint64_t db = smem_desc(sb);
        int ls = 0;
        uint32_t par = 0;
        for (int t = 0; t < tiles; ++t) {
            uint64_t *bar = &bars[warp * DEPTH + ls];
            if (t >= DEPTH) {
                if (ONE_LANE_WAIT) { if (lane == 0) mbar_wait(bar, par ^ 1u); __syncwarp(); }
                else mbar_wait(bar, par ^ 1u);
            }
            tc_fence_after();
            if (elect_one()) {
                const uint32_t d = tb + 32 * NI + (warp * DEPTH + ls) * 64;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ts(d, tb + warp * 32 + k * 8, db + 2 * k, idesc, k > 0);
                umma_commit(bar);
            }
            __syncwarp();
            if (++ls == DEPTH) { ls = 0; par ^= 1u; }
        }
        // drain
        for (int i = 0; i < DEPTH && i < tiles; ++i) {
            mbar_wait(&bars[warp * DEPTH + ls], par ^ 1u);
            if (++ls == DEPTH) { ls = 0; par ^= 1u; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tb, 512);
}

template <int NI, int DEPTH, bool ONE_LANE_WAIT>
void run_mma2(const char *name, int sms) {
    const int tiles = 40000;
    const size_t smem = 8192 + 256;
    auto kern = mma_probe2<NI, DEPTH, ONE_LANE_WAIT>;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    kern<<<sms, 128, smem>>>(2000, 1u);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    kern<<<sms, 128, smem>>>(tiles, 7u);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double mmas = 4.0 * NI * tiles;
    printf("mma2 %-40s %8.3f ms  %7.1f ns per tile, %6.1f clk per MMA @1.965 (chip-wide: %d issuers)\n", name, ms, ms * 1e6 / tiles,
           ms * 1e6 / mmas * 1.965, NI);
    fflush(stdout);
}

// ------------------------------------------------------------------------------------------------------
// (3) TS-form correctness, fp16 accumulators, pack::16b layout
// ------------------------------------------------------------------------------------------------------
// D[128 x 64] = A[128 x 64] * B[64 x 64]^T.  A is written to tensor memory by its row's thread (lane = row, column c
// holds elements 2c, 2c+1); B sits in shared memory K-major with the 128-byte swizzle.  mode bit 0: fp16 operands
// (else bf16); bit 1: fp16 accumulators read back with pack::16b; bit 2: SS form (A from shared memory) as control.
__global__ void __launch_bounds__(128, 1) ts_check(const uint16_t *A, const uint16_t *B, float *D, int mode) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char *sa = smem;            // 16 KB (SS control)
    unsigned char *sb = smem + 16384;    // 8 KB
    uint64_t *bar = reinterpret_cast<uint64_t *>(smem + 16384 + 8192);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(bar + 1);
    const int tid = threadIdx.x;
    for (int i = tid; i < 64 * 64; i += 128) {
        const int n = i / 64, k = i % 64;
        const int off = n * 128 + ((((k * 2) / 16) ^ (n % 8)) * 16) + (k * 2) % 16;
        *reinterpret_cast<uint16_t *>(sb + off) = B[i];
    }
    for (int i = tid; i < 128 * 64; i += 128) {
        const int m = i / 64, k = i % 64;
        const int off = m * 128 + ((((k * 2) / 16) ^ (m % 8)) * 16) + (k * 2) % 16;
        *reinterpret_cast<uint16_t *>(sa + off) = A[i];
    }
    if (tid == 0) {
        mbar_init(bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (tid < 32) tmem_alloc(tmem_slot, 256);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tb = *tmem_slot;
    const uint32_t lane_base = tb + (static_cast<uint32_t>((tid >> 5) * 32) << 16);
    // A row -> tensor memory columns [0, 32)
    uint32_t av[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) av[c] = reinterpret_cast<const uint32_t *>(A + tid * 64)[c];
    tmem_st32(lane_base, av);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int ab_fmt = (mode & 1) ? 0 : 1;
    const int c_fmt = (mode & 2) ? 0 : 1;
    const uint32_t idesc = make_idesc(64, c_fmt, ab_fmt);
    if (tid == 0) {
        const uint64_t db = smem_desc(sb), da = smem_desc(sa);
        for (int k = 0; k < 4; ++k) {
            if (mode & 4)
                umma_ss(tb + 64, da + 2 * k, db + 2 * k, idesc, k > 0);
            else
                umma_ts(tb + 64, tb + k * 8, db + 2 * k, idesc, k > 0);
        }
        umma_commit(bar);
    }
    mbar_wait(bar, 0);
    tc_fence_after();
    uint32_t v[32], w[32];
    if (mode & 2) {
        tmem_ld64_pack16(lane_base + 64, v);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) {
            const __half2 h = *reinterpret_cast<__half2 *>(&v[c]);
            D[tid * 64 + 2 * c] = __low2float(h);
            D[tid * 64 + 2 * c + 1] = __high2float(h);
        }
    } else {
        tmem_ld32(lane_base + 64, v);
        tmem_ld32(lane_base + 96, w);
        tmem_ld_wait();
        for (int c = 0; c < 32; ++c) {
            D[tid * 64 + c] = __uint_as_float(v[c]);
            D[tid * 64 + 32 + c] = __uint_as_float(w[c]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (tid < 32) tmem_dealloc(tb, 256);
}

static uint16_t f2bf(float f) {
    uint32_t u;
    memcpy(&u, &f, 4);
    u += 0x7FFFu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
}
static float bf2f(uint16_t h) {
    uint32_t u = static_cast<uint32_t>(h) << 16;
    float f;
    memcpy(&f, &u, 4);
    return f;
}

void run_ts_check(int mode) {
    std::vector<uint16_t> A(128 * 64), B(64 * 64);
    std::vector<float> Af(128 * 64), Bf(64 * 64);
    uint32_t s = 12345u + mode;
    auto rnd = [&]() {
        s = s * 1664525u + 1013904223u;
        return (static_cast<float>((s >> 8) & 0xFFFF) / 65536.0f - 0.5f) * 2.0f;
    };
    for (int i = 0; i < 128 * 64; ++i) {
        const float x = rnd();
        if (mode & 1) {
            const __half h = __float2half_rn(x);
            A[i] = *reinterpret_cast<const uint16_t *>(&h);
            Af[i] = __half2float(h);
        } else {
            A[i] = f2bf(x);
            Af[i] = bf2f(A[i]);
        }
    }
    for (int i = 0; i < 64 * 64; ++i) {
        const float x = rnd();
        if (mode & 1) {
            const __half h = __float2half_rn(x);
            B[i] = *reinterpret_cast<const uint16_t *>(&h);
            Bf[i] = __half2float(h);
        } else {
            B[i] = f2bf(x);
            Bf[i] = bf2f(B[i]);
        }
    }
    uint16_t *dA, *dB;
    float *dD;
    CK(cudaMalloc(&dA, A.size() * 2));
    CK(cudaMalloc(&dB, B.size() * 2));
    CK(cudaMalloc(&dD, 128 * 64 * 4));
    CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dD, 0, 128 * 64 * 4));
    const size_t smem = 16384 + 8192 + 64;
    ts_check<<<1, 128, smem>>>(dA, dB, dD, mode);
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * 64);
    CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0;
    int n_rn = 0, n_rz = 0, n_other = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < 64; ++n) {
            double ref = 0;
            for (int k = 0; k < 64; ++k) ref += static_cast<double>(Af[m * 64 + k]) * Bf[n * 64 + k];
            const double got = D[m * 64 + n];
            max_err = fmax(max_err, fabs(got - ref));
            max_ref = fmax(max_ref, fabs(ref));
            if (mode & 2) {
                // how does the final value relate to the exact sum?  (only indicative: the accumulator is rounded
                // after each of the 4 K-steps)
                const __half rn = __float2half_rn(static_cast<float>(ref));
                const __half rz = __float2half_rz(static_cast<float>(ref));
                if (got == __half2float(rn)) ++n_rn;
                else if (got == __half2float(rz)) ++n_rz;
                else ++n_other;
            }
        }
    printf("ts_check mode %d (%s operands, %s accumulators, %s): max |err| %.3e  (max |ref| %.3f)", mode, (mode & 1) ? "f16" : "bf16",
           (mode & 2) ? "f16" : "f32", (mode & 4) ? "SS" : "TS", max_err, max_ref);
    if (mode & 2) printf("  final==RN(exact) %d, ==RZ(exact) %d, other %d", n_rn, n_rz, n_other);
    printf("\n");
    printf("   D[0][0..3] = %g %g %g %g   D[5][60..63] = %g %g %g %g\n", D[0], D[1], D[2], D[3], D[5 * 64 + 60], D[5 * 64 + 61],
           D[5 * 64 + 62], D[5 * 64 + 63]);
    cudaFree(dA);
    cudaFree(dB);
    cudaFree(dD);
}

int main(int argc, char **argv) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    printf("SMs %d\n", sms);
    const char *what = argc > 1 ? argv[1] : "all";
    if (!strcmp(what, "ts")) {
        run_ts_check(argc > 2 ? atoi(argv[2]) : 0);
        return 0;
    }

    if (!strcmp(what, "mma")) goto mma_only;
    {
    uint32_t *in, *out;
    CK(cudaMalloc(&in, 8192 * 4));
    CK(cudaMalloc(&out, static_cast<size_t>(sms) * 512 * 4));
    std::vector<uint32_t> h(8192);
    for (int i = 0; i < 8192; ++i) h[i] = 0x3C003C00u + i * 7u;
    CK(cudaMemcpy(in, h.data(), 8192 * 4, cudaMemcpyHostToDevice));
    run_alu<4>("FFMA", in, out, sms, 1);
    run_alu<0>("FMNMX3", in, out, sms, 1);
    run_alu<1>("VHMNMX (f16x2 max3)", in, out, sms, 1);
    run_alu<2>("VIMNMX3.S16x2", in, out, sms, 1);
    run_alu<3>("LOP3", in, out, sms, 1);
    run_alu<5>("HSETP2+SEL+IADD", in, out, sms, 1);
    run_alu<7>("FSETP+@P IADD", in, out, sms, 1);
    run_alu<6>("FMNMX3 | FFMA alternating", in, out, sms, 1);
    run_alu<8>("VHMNMX | FFMA alternating", in, out, sms, 1);
    }
mma_only:

    if (!strcmp(what, "all") || !strcmp(what, "mma")) {
        const int which = argc > 2 ? atoi(argv[2]) : -1;
        int id = 0;
#define RUN(...)                                    \
    do {                                            \
        if (which < 0 || which == id) __VA_ARGS__;  \
        ++id;                                       \
    } while (0)
        RUN(run_mma<false, 64, 1, 0, 0>("SS N=64 f32", sms));
        RUN(run_mma<false, 64, 1, 1, 0>("SS N=64 f32", sms));
        RUN(run_mma<false, 64, 1, 1, 1>("SS N=64 f32", sms));
        RUN(run_mma<false, 64, 1, 2, 0>("SS N=64 f32", sms));
        RUN(run_mma<true, 64, 1, 1, 0>("TS N=64 f32", sms));
        RUN(run_mma<true, 64, 1, 1, 1>("TS N=64 f32", sms));
        RUN(run_mma<true, 64, 1, 2, 0>("TS N=64 f32", sms));
        RUN(run_mma<true, 64, 0, 1, 0>("TS N=64 f16 operands+acc", sms));
        RUN(run_mma<false, 128, 1, 1, 0>("SS N=128 f32", sms));
        RUN(run_mma<false, 128, 1, 1, 1>("SS N=128 f32", sms));
        RUN(run_mma<true, 128, 1, 1, 0>("TS N=128 f32", sms));
        RUN(run_mma<false, 256, 1, 1, 0>("SS N=256 f32", sms));
        RUN(run_mma<true, 192, 1, 1, 0>("TS N=192 f32", sms));
        RUN((run_mma2<2, 3, false>("2 issuers, 3 slots, commit per 4 MMAs", sms)));
        RUN((run_mma2<2, 3, true>("2 issuers, 3 slots, lane-0 wait", sms)));
        RUN((run_mma2<3, 2, false>("3 issuers, 2 slots", sms)));
        RUN((run_mma2<1, 3, false>("1 issuer, 3 slots", sms)));
        RUN((run_mma2<4, 1, false>("4 issuers, 1 slot each (fully serial per issuer)", sms)));
    }
    printf("done\n");
    return 0;
}
