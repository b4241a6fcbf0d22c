// K2 tensor engine for d <= 64 (full-ranking evaluation, Model/LightGCN.py:147-155; also the 64-d item
// graphs of Model/LATTICE.py:46-52): fp16 tcgen05 pre-filter with a fused threshold epilogue, exact fp32
// re-score and certification.  Replaces the round-1 `tc_candidates_kernel<true>`.
//
// What changed against round 1, and the measurement behind each change (profiles/r02_probe_*.log):
//   * A operand in TENSOR MEMORY (tcgen05.mma "TS" form).  A 128x64x16 MMA with both operands in shared memory
//     reads 6 KB per instruction and is bound by the 128 B/clk shared-memory port at ~50 clk; with the 512 user
//     rows of a unit parked in 128 TMEM columns (written once per unit by the rows' own epilogue threads with
//     tcgen05.st) only the 2 KB item slab is read: 34 clk per MMA = the nominal tensor peak.
//   * TWO issuing warps, warp-uniform control flow with an ELECTed lane.  `if (lane == 0)` around tcgen05.mma
//     makes the compiler wrap every MMA in an R2UR/BRA.U.ANY loop (72 clk per MMA); one converged warp still
//     needs ~50 clk between dispatches, two warps (row blocks 0-1 / 2-3) reach 34.
//   * SIX accumulator slots of 128 rows x 64 columns with their own full/empty barriers instead of two
//     512-row stages: a row block is handed back as soon as its four warps have read it.
//   * fp16 operands, each user row and the item table scaled by a power of two (exact): 2^-11 relative operand
//     error instead of bf16's 2^-8, so the certification margin eps shrinks four-fold.
//   * The epilogue keeps candidates only: per 64 scores, eight 8-column maxima (FMNMX3 trees) and one test;
//     a warp looks at individual columns only in 8-column groups where one of its rows beats its threshold.
//     No sort, no mask look-ups, no list hand-over inside the MMA kernel: lists live in their final place
//     in global memory and the unit's last act is two 4-byte stores per row.  Sorting, mask removal and the
//     top-K' cut moved into the re-score kernel (a plain SIMT kernel that is not on the tensor pipeline).
//   * The sample pre-pass is the same kernel in SEED mode over every `stride`-th column tile: the epilogue
//     keeps the 16 largest TILE MAXIMA of each row in registers (a branch-free insertion network) -- the k-th
//     largest tile maximum is a lower bound of the k-th largest sample score -- and never touches memory.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cmath>
#include <cstdlib>

#include "common.cuh"
#include "engine.cuh"
#include "tc_ptx.cuh"

// Diagnostic builds only (scripts/gpu_eval_floor.sh): 1 = maxima and threshold test but nothing is listed, 2 = accumulators
// are read and dropped, 3 = accumulators are not even read.  Results are wrong (every row falls back) -- timing only.
#ifndef CHAOREC_EV_DEBUG
#define CHAOREC_EV_DEBUG 0
#endif
// Waiting style (measured, scripts/gpu_eval_floor.sh): bit 0 = the MMA issuers poll their barriers, bit 1 = the epilogue
// warps poll the accumulator-full barrier, before falling back to the suspended bounded wait.
#ifndef CHAOREC_EV_SPIN
#define CHAOREC_EV_SPIN 0
#endif

namespace chaorec {
namespace {
using namespace ptx;

constexpr int EV_BM = 384;                        // user rows per unit (3 MMA row blocks)
constexpr int EV_BN = 64;                         // item columns per tile
constexpr int EV_RB = 3;
constexpr int EV_KPAD = 64;                       // inner dimension of the fp16 copies (d zero-padded)
constexpr int EV_STAGES = 8;                      // item-slab ring
constexpr int EV_STAGE_BYTES = EV_BN * EV_KPAD * 2;
constexpr int EV_SLOTS = 2 * EV_RB;               // accumulator slots (128 lanes x 64 columns): two private ones per row block
constexpr int EV_ACC0 = 32 * EV_RB;               // TMEM columns [0,96) hold A: row block rb at 32*rb
constexpr int EV_GCAP = 160;                      // 8-column group records per (row, slice, column half)
constexpr int EV_REC_WORDS = 12;                  // a record: 8 raw scores, first column, 3 words of padding (16-byte aligned stores)
constexpr int EV_REC_BYTES = EV_REC_WORDS * 4;
constexpr int EV_GLIMIT = EV_GCAP - 4;            // (a half-tile visit appends at most 4 records)
constexpr int EV_MASK_CAP = 256;                  // train items of a row that the approximate cut makes room for
constexpr int EV_CH = 2;                          // epilogue threads per row: each takes 32 of a tile's 64 columns
constexpr int EV_EPI_WARPS = 4 * EV_RB * EV_CH;   // warp w: row block w / 8, column half (w / 4) % 2, TMEM lane quarter w % 4
constexpr int EV_TMA_WARP = EV_EPI_WARPS;         // warp 24
#ifndef CHAOREC_EV_ISSUERS
#define CHAOREC_EV_ISSUERS 2
#endif
constexpr int EV_ISS = CHAOREC_EV_ISSUERS;        // issuing warps per row block (2: one per accumulator slot)
constexpr int EV_MMA_WARP0 = EV_EPI_WARPS + 1;    // warps 25 .. : issuer w serves row block w / EV_ISS, tiles with parity w % EV_ISS
constexpr int EV_THREADS = (EV_EPI_WARPS + 1 + EV_RB * EV_ISS) * 32;   // 992 threads -> 64 registers each
constexpr int EV_SEEDK = 16;                      // tile maxima kept per row in seed mode
constexpr int EV_SEED_TAB = 128;                  // per-row seed rank is looked up by the row's train-item count
constexpr int EV_NUM_BARS = 2 * EV_STAGES + 2 * EV_SLOTS + EV_RB;
constexpr uint32_t EV_IDESC = make_idesc_m128(EV_BN, /*f32 acc*/ 1, /*f16 operands*/ 0);
constexpr int EV_KP_MAX = 128;

struct EvalArgs {
    int64_t Ma, Nb;
    int n_row_tiles, n_slices, tiles_per_slice, n_col_tiles;
    int n_seed_tiles, seed_stride, seed_off;           // seed mode: tile i of a unit = seed_off + i * seed_stride
    unsigned char seed_k[EV_SEED_TAB];                 // seed rank (1-based, <= EV_SEEDK) by train-item count of the row
    const __half *a16;            // [Ma][64] scaled fp16 rows
    const int64_t *mask_rowptr;   // only the row lengths are used here
    int exclude_self;
    int kp;
    uint32_t *recs;               // [Ma][n_slices][EV_CH][EV_GCAP] group records (EV_REC_WORDS words each), 16-byte aligned
    int32_t *cnt_out;             // [Ma][n_slices][EV_CH] records written
    float *tau_out;               // [Ma][n_slices][EV_CH] threshold in force at the end of the unit (+inf: row gave up)
    float *seed_scratch;          // seed mode: [gridDim.x][EV_BM][EV_CH][EV_SEEDK] half-tile maxima of the CTA's current unit
    float *tau0;                  // [Ma] starting thresholds: written in seed mode, read in main mode (nullable)
    unsigned int *dbg;            // [0] = pipeline time-out flag, [1..7] = first time-out record (see ev_wait)
};

// Every mbarrier wait of the pipeline is bounded: a protocol error (or a lost TMA / MMA completion) must end in a
// reported failure, never in a hung GPU.  The fast path is one try_wait; only a wait that has to retry reads the
// clock and the CTA's abort flag.  On a time-out the CTA records which wait failed (role/kind/index/parity, the
// work position and the barrier word), raises the flag -- every later wait of the CTA then returns at once and the
// kernel runs to completion on garbage -- and the re-score kernel sends ALL rows to the exact engine, so results
// stay correct; `dbg[0] != 0` is surfaced through the host shim's stats.
constexpr unsigned long long EV_WAIT_NS = 4000000000ull;   // 4 s
// Debug builds (-DCHAOREC_EV_TRACE): every warp keeps a (position << 8 | marker) word in shared memory; the first
// time-out of a CTA dumps all of them after the record.
#ifdef CHAOREC_EV_TRACE
#define EV_TRACE(marker, posv) do { if (lane == 0) s_trace[warp] = (static_cast<uint32_t>(posv) << 8) | (marker); } while (0)
#else
#define EV_TRACE(marker, posv) do { } while (0)
#endif
// try_wait with a suspend-time hint: a waiting warp sleeps in hardware until the phase completes (or the hint
// expires) instead of polling through the issue slots of the warps that have work (with plain try_wait the 16
// epilogue warps spent a quarter of the kernel's instructions polling: profiles/r02_eval_v2a_*).
__device__ __forceinline__ bool mbar_try_wait_hint(uint64_t *bar, uint32_t parity, uint32_t ns) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"(ns)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void ev_wait(uint64_t *bar, uint32_t parity, volatile uint32_t *s_abort, unsigned int *dbg, uint32_t code,
                                        uint32_t pos) {
    if (mbar_try_wait_hint(bar, parity, 20000u)) return;
    uint32_t n = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait_hint(bar, parity, 20000u)) {
        if ((++n & 63u) != 0u) continue;                   // look at the clock and the abort flag every 64th retry
        if (*s_abort) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t0 == 0) t0 = t1;
        if (t1 - t0 > EV_WAIT_NS) {
            *s_abort = 1u;
            if (dbg != nullptr && atomicExch(&dbg[0], 1u) == 0u) {
                dbg[1] = code;
                dbg[2] = pos;
                dbg[3] = blockIdx.x;
                dbg[4] = parity;
                const unsigned long long w = *reinterpret_cast<volatile unsigned long long *>(bar);
                dbg[5] = static_cast<unsigned int>(w);
                dbg[6] = static_cast<unsigned int>(w >> 32);
                dbg[7] = threadIdx.x;
#ifdef CHAOREC_EV_TRACE
                for (int w2 = 0; w2 < 20; ++w2) dbg[8 + w2] = s_abort[3 + w2];
#endif
            }
            return;
        }
    }
}

// Epilogue flavour of the bounded wait, on a precomputed shared-memory address: the common case (phase complete) is a
// single try_wait and a branch.  The out-of-line retry loop takes three registers of arguments only (the abort flag and
// the diagnostics pointer live at fixed shared-memory offsets), so the call site does not drag address arithmetic
// into the tile loop.
__device__ __forceinline__ bool mbar_try_wait_a(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return ok != 0;
}
// Non-blocking probe: a warp that polls with it notices a completed phase within a few tens of clocks, where a
// suspended try_wait takes a few hundred to wake up (the accumulator hand-over is a latency loop: tensor memory holds
// only ~800 clocks of MMA work at d = 64).
__device__ __forceinline__ bool mbar_test_wait_a(uint32_t bar_addr, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_spin_a(uint32_t bar_addr, uint32_t parity, int polls) {
#pragma unroll 1
    for (int i = 0; i < polls; ++i)
        if (mbar_test_wait_a(bar_addr, parity)) return true;
    return false;
}
constexpr int EV_CTRL_OFF = EV_STAGES * EV_STAGE_BYTES + EV_NUM_BARS * 8;   // control words behind the barriers
struct EvCtrl {
    uint32_t tmem_base;
    uint32_t abort_flag;
    unsigned int *dbg;
    uint32_t trace[28];
};
__device__ __noinline__ void ev_wait_slow_a(uint32_t bar_addr, uint32_t parity, uint32_t code_pos) {
    extern __shared__ __align__(1024) unsigned char ev_smem_w[];
    volatile EvCtrl *ctrl = reinterpret_cast<volatile EvCtrl *>(ev_smem_w + EV_CTRL_OFF);
    uint32_t n = 0;
    unsigned long long t0 = 0;
    while (!mbar_try_wait_a(bar_addr, parity)) {
        if ((++n & 63u) != 0u) continue;
        if (ctrl->abort_flag) return;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t0 == 0) t0 = t1;
        if (t1 - t0 > EV_WAIT_NS) {
            ctrl->abort_flag = 1u;
            unsigned int *dbg = ctrl->dbg;
            if (dbg != nullptr && atomicExch(&dbg[0], 1u) == 0u) {
                dbg[1] = code_pos >> 20;
                dbg[2] = code_pos & 0xFFFFFu;
                dbg[3] = blockIdx.x;
                dbg[4] = parity;
                dbg[7] = threadIdx.x;
            }
            return;
        }
    }
}

__device__ __forceinline__ float max8(const uint32_t *v) {
    float m = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
    for (int j = 2; j < 8; ++j) m = fmaxf(m, __uint_as_float(v[j]));
    return m;
}
__device__ __forceinline__ float max8f(const float (&v)[8]) {
    float m = fmaxf(v[0], v[1]);
#pragma unroll
    for (int j = 2; j < 8; ++j) m = fmaxf(m, v[j]);
    return m;
}

// Branch-free conditional store of one 8-column group (raw scores + first column) at the end of the row's record
// list in global memory: setp + three predicated stores + a predicated pointer bump.  Which of the eight scores
// beat the threshold is decided by the re-score kernel; the MMA kernel only decides "this group has one".
__device__ __forceinline__ void store_group_if_above(uint32_t *&p, float m, float tau, const uint32_t *v, uint32_t col) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.gt.f32 p, %1, %2;\n\t"
        "@p st.global.v4.u32 [%0], {%3, %4, %5, %6};\n\t"
        "@p st.global.v4.u32 [%0+16], {%7, %8, %9, %10};\n\t"
        "@p st.global.u32 [%0+32], %11;\n\t"
        "@p add.u64 %0, %0, 48;\n\t"
        "}"
        : "+l"(p)
        : "f"(m), "f"(tau), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(col)
        : "memory");
}

__device__ __forceinline__ uint64_t key_of_raw(uint64_t raw) {
    return make_key(__uint_as_float(static_cast<uint32_t>(raw >> 32)), static_cast<uint32_t>(raw & 0xFFFFFFFFull));
}
__device__ __forceinline__ uint64_t raw_of_key(uint64_t key) {
    return (static_cast<uint64_t>(__float_as_uint(key_score(key))) << 32) | key_index(key);
}

// In a 256-bin histogram (bins in DESCENDING order of value) find the bin in which the running count from
// the top reaches `rank1` (1-based) and the count strictly above that bin.
__device__ __forceinline__ void hist_find(const uint32_t *hist, int lane, int rank1, int &bin, int &above) {
    uint32_t my[8];
    int sum = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {          // lane l owns bins 255-8l .. 248-8l
        my[j] = hist[255 - 8 * lane - j];
        sum += static_cast<int>(my[j]);
    }
    int incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    const int excl = incl - sum;
    const bool has = excl < rank1 && rank1 <= incl;
    int b = 0, ab = 0;
    if (has) {
        int cum = excl;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (b == 0 && cum + static_cast<int>(my[j]) >= rank1) { b = 256 - 8 * lane - j; ab = cum; }   // b stored +1
            cum += static_cast<int>(my[j]);
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, has);
    const int src = m ? (__ffs(m) - 1) : 0;
    bin = __shfl_sync(0xffffffffu, b, src) - 1;
    above = __shfl_sync(0xffffffffu, ab, src);
    if (m == 0u) { bin = 0; above = 0; }
}

// Rank selection over n orderable keys (global or shared memory; 0 = empty slot), executed by a whole warp:
// returns a 32-bit prefix P (a lower bound of the (r+1)-th largest key's score bits: the 16 bits below the
// highest bit in which the keys differ are resolved by two 256-bin histogram passes) such that at least r+1
// keys have score bits >= P whenever n > r.  Cheap replacement for a sort when only a cut is needed.
template <typename LoadKey>
__device__ __forceinline__ uint32_t select_prefix(LoadKey key_at, int n, int r, int lane, uint32_t *hist) {
    uint32_t lo = 0xFFFFFFFFu, mx = 0u;
    for (int i = lane; i < n; i += 32) {
        const uint64_t k = key_at(i);
        if (k != 0ull) {
            const uint32_t hi = static_cast<uint32_t>(k >> 32);
            lo = min(lo, hi);
            mx = max(mx, hi);
        }
    }
    lo = __reduce_min_sync(0xffffffffu, lo);
    mx = __reduce_max_sync(0xffffffffu, mx);
    uint32_t prefix = mx;
    if (lo < mx) {
        const int p = 31 - __clz(lo ^ mx);                 // highest differing bit
        const int s1 = p >= 7 ? p - 7 : 0;
        const uint32_t common = (s1 + 8 >= 32) ? 0u : ((mx >> (s1 + 8)) << (s1 + 8));
#pragma unroll
        for (int j = 0; j < 8; ++j) hist[lane + 32 * j] = 0u;
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            const uint64_t k = key_at(i);
            if (k != 0ull) atomicAdd(&hist[(static_cast<uint32_t>(k >> 32) >> s1) & 255u], 1u);
        }
        __syncwarp();
        int bin1, above1;
        hist_find(hist, lane, r + 1, bin1, above1);
        prefix = common | (static_cast<uint32_t>(bin1) << s1);
        if (s1 > 0) {
            const int s2 = s1 >= 8 ? s1 - 8 : 0;
            const uint32_t sel = (common >> s1) | static_cast<uint32_t>(bin1);
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 8; ++j) hist[lane + 32 * j] = 0u;
            __syncwarp();
            for (int i = lane; i < n; i += 32) {
                const uint64_t k = key_at(i);
                const uint32_t hi = static_cast<uint32_t>(k >> 32);
                if (k != 0ull && (hi >> s1) == sel) atomicAdd(&hist[(hi >> s2) & ((1u << (s1 - s2)) - 1u)], 1u);
            }
            __syncwarp();
            int bin2, above2;
            hist_find(hist, lane, r + 1 - above1, bin2, above2);
            prefix |= static_cast<uint32_t>(bin2) << s2;
        }
        __syncwarp();
    }
    return prefix;
}

template <bool SEED>
__global__ void __launch_bounds__(EV_THREADS, 1)
tc_eval_kernel(const __grid_constant__ CUtensorMap map_b, const EvalArgs a) {
    extern __shared__ __align__(1024) unsigned char ev_smem[];
    unsigned char *ring = ev_smem;
    uint64_t *bars = reinterpret_cast<uint64_t *>(ev_smem + EV_STAGES * EV_STAGE_BYTES);
    uint64_t *full_bar = bars;                              // [EV_STAGES] item slab landed
    uint64_t *empty_bar = bars + EV_STAGES;                 // [EV_STAGES] all issuers' MMAs on the slab retired
    // Every row block owns two accumulator slots (2*rb, 2*rb+1) and alternates between them: tile g of the CTA
    // uses slot 2*rb + (g & 1), barrier phase g >> 1.  (A first version shared six slots between four row blocks;
    // a warp could then be waiting for its row block's NEXT use of a slot while the other row block's use in
    // between had not happened yet -- one parity bit cannot tell those apart -- and, worse, a slot was held until
    // the slowest warp of ANOTHER row block had read it, which serialised MMA and filter.)
    uint64_t *tfull_bar = bars + 2 * EV_STAGES;             // [EV_SLOTS]  accumulator slot complete
    uint64_t *tempty_bar = tfull_bar + EV_SLOTS;            // [EV_SLOTS]  the eight warps of the row block have read it
    uint64_t *afull_bar = tempty_bar + EV_SLOTS;            // [EV_RB]     A rows of the unit are in tensor memory
    EvCtrl *ctrl = reinterpret_cast<EvCtrl *>(ev_smem + EV_CTRL_OFF);
    volatile uint32_t *s_abort = &ctrl->abort_flag;
    volatile uint32_t *s_trace = ctrl->trace;          // (debug builds)
    (void)s_trace;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        ctrl->abort_flag = 0u;
        ctrl->dbg = a.dbg;
        for (int s = 0; s < EV_STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], EV_RB);
        }
        for (int s = 0; s < EV_SLOTS; ++s) {
            mbar_init(&tfull_bar[s], 1);
            mbar_init(&tempty_bar[s], 4 * EV_CH);
        }
        for (int s = 0; s < EV_RB; ++s) mbar_init(&afull_bar[s], 4 * EV_CH);
#ifdef CHAOREC_EV_XCOMMIT
        for (int s = 0; s < EV_RB; ++s) mbar_init(reinterpret_cast<uint64_t *>(ctrl->trace + 4 + 2 * s), 1);
#endif
        mbar_fence_init();
    }
    if (warp == EV_MMA_WARP0) tmem_alloc(&ctrl->tmem_base, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = ctrl->tmem_base;
    const int n_units = a.n_row_tiles * a.n_slices;
    if (warp == EV_TMA_WARP) {
        // ================= TMA producer: one 64-item x 64-element fp16 slab per tile =================
        int stage = 0;
        uint32_t phase = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int sl = unit % a.n_slices;
            const int t_lo = sl * a.tiles_per_slice;
            const int t_hi = min(t_lo + a.tiles_per_slice, a.n_col_tiles);
            const int n_t = SEED ? a.n_seed_tiles : (t_hi - t_lo);
            for (int i = 0; i < n_t; ++i) {
                const int t = SEED ? (a.seed_off + i * a.seed_stride) : (t_lo + i);
                EV_TRACE(0x61, i);
                ev_wait(&empty_bar[stage], phase ^ 1u, s_abort, a.dbg, 0x100u | stage, static_cast<uint32_t>(i));
                __syncwarp();
                if (elect_one()) {
                    mbar_arrive_expect_tx(&full_bar[stage], EV_STAGE_BYTES);
                    tma_load_2d(ring + stage * EV_STAGE_BYTES, &map_b, 0, t * EV_BN, &full_bar[stage]);
                }
                __syncwarp();
                if (++stage == EV_STAGES) { stage = 0; phase ^= 1u; }
            }
        }
    } else if (warp >= EV_MMA_WARP0) {
        // ================= MMA issuers =================
        // A converged warp dispatches one 128x64x16 tcgen05.mma every ~50 clocks (profiles/r02_probe_mma_issue_rates.log),
        // and a tile costs it two barrier probes, four MMAs and two commits: ~500 clocks per tile, which was the floor
        // of the whole kernel with one issuer per row block.  With two issuers per row block each owns one of the
        // block's two accumulator slots and takes every other tile.
        const int rb = (warp - EV_MMA_WARP0) / EV_ISS;
        const int par = (warp - EV_MMA_WARP0) % EV_ISS;
        uint32_t uphase = 0;
        int g0 = 0;                                          // tiles of the CTA before the current unit
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int sl = unit % a.n_slices;
            const int t_lo = sl * a.tiles_per_slice;
            const int t_hi = min(t_lo + a.tiles_per_slice, a.n_col_tiles);
            const int n_t = SEED ? a.n_seed_tiles : (t_hi - t_lo);
            ev_wait(&afull_bar[rb], uphase, s_abort, a.dbg, 0x200u | rb, static_cast<uint32_t>(unit));
            uphase ^= 1u;
            for (int i = par; i < n_t; i += EV_ISS) {        // (tile counts are even: the parity of g is the parity of i)
                const int g = g0 + i;
                const int slot = 2 * rb + (g & 1);
                const int stage = g % EV_STAGES;
                const uint32_t phase = static_cast<uint32_t>((g / EV_STAGES) & 1);
                EV_TRACE(0x51, i);
                // both probes are in flight before either result is needed (a completed try_wait still costs ~100 clk)
                const uint32_t par_e = static_cast<uint32_t>(((g >> 1) & 1) ^ 1);
#if CHAOREC_EV_SPIN & 1
                const bool ok_e = mbar_spin_a(smem_u32(&tempty_bar[slot]), par_e, 256);
                const bool ok_f = mbar_spin_a(smem_u32(&full_bar[stage]), phase, 256);
#else
                const bool ok_e = mbar_try_wait(&tempty_bar[slot], par_e);
                const bool ok_f = mbar_try_wait(&full_bar[stage], phase);
#endif
                if (!ok_e) ev_wait(&tempty_bar[slot], par_e, s_abort, a.dbg, 0x300u | slot, static_cast<uint32_t>(i));
                EV_TRACE(0x52, i);
                if (!ok_f) ev_wait(&full_bar[stage], phase, s_abort, a.dbg, 0x400u | (rb << 4) | stage, static_cast<uint32_t>(i));
                EV_TRACE(0x53, i);
                tc_fence_after();
                __syncwarp();
                if (elect_one()) {
                    const uint64_t db = smem_desc_sw128(ring + stage * EV_STAGE_BYTES);
                    const uint32_t d0 = tmem_base + EV_ACC0 + slot * EV_BN;
                    const uint32_t a0 = tmem_base + rb * 32;
#pragma unroll
                    for (int k = 0; k < 4; ++k) umma_ts(d0, a0 + 8 * k, db + 2 * k, EV_IDESC, k > 0 ? 1u : 0u);
                    umma_commit(&tfull_bar[slot]);
                    umma_commit(&empty_bar[stage]);
#ifdef CHAOREC_EV_XCOMMIT    // (timing experiment: what does one more tcgen05.commit per tile cost?)
                    for (int x = 0; x < CHAOREC_EV_XCOMMIT; ++x) umma_commit(reinterpret_cast<uint64_t *>(ctrl->trace + 4 + 2 * rb));
#endif
                }
                __syncwarp();
                EV_TRACE(0x54, i);
            }
            g0 += n_t;
        }
    } else {
        // ================= epilogue: two threads per user row, 32 columns of every tile each =================
        // (One thread per row and tile leaves 3 warps per scheduler, each a long dependent chain -- barrier wait,
        // TMEM load, max tree, vote -- of ~750 clk per tile: the profile showed the issue ports half idle with the
        // tensor pipe waiting.  Halving the columns per thread doubles the warps that hide each other's latency.)
        const int rb = warp >> 3;                // 128-row block of the unit
        const int h = (warp >> 2) & 1;           // column half of every tile
        const int q = warp & 3;                  // TMEM lane quarter this warp may access
        const int row_local = rb * 128 + q * 32 + lane;
        const uint32_t lane_base = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
        // tile g of the CTA uses slot 2*rb + (g & 1), phase g >> 1; units have even tile counts.  The three addresses
        // are made opaque to the compiler: left visible, it re-derives them from %tid inside the tile loop (17
        // instructions per visit in the round-2a build) instead of keeping three registers alive.
        uint32_t bar_full0 = smem_u32(&tfull_bar[2 * rb]), bar_empty0 = smem_u32(&tempty_bar[2 * rb]);
        uint32_t taddr0 = lane_base + EV_ACC0 + (2 * rb) * EV_BN + h * 32;
        uint32_t lane0 = lane == 0 ? 1u : 0u;
        asm volatile("" : "+r"(bar_full0), "+r"(bar_empty0), "+r"(taddr0), "+r"(lane0));
        uint32_t spar = 0;
        for (int unit = blockIdx.x; unit < n_units; unit += gridDim.x) {
            const int rt = unit / a.n_slices, sl = unit % a.n_slices;
            const int t_lo = sl * a.tiles_per_slice;
            const int t_hi = min(t_lo + a.tiles_per_slice, a.n_col_tiles);
            const int n_t = SEED ? a.n_seed_tiles : (t_hi - t_lo);
            const int64_t row = static_cast<int64_t>(rt) * EV_BM + row_local;
            const bool row_ok = row < a.Ma;
            // ---- this thread's half of its user row -> tensor memory (16 of the columns 32*rb .. 32*rb+31 of its
            // lane).  Every MMA of the previous unit that read these columns has retired: this warp has seen the
            // accumulator of the unit's last tile of its row block, and a commit covers all MMAs issued before it.
            {
                uint32_t av[16];
                if (row_ok) {
                    const uint4 *src = reinterpret_cast<const uint4 *>(a.a16 + row * EV_KPAD + h * 32);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 x = src[j];
                        av[4 * j] = x.x; av[4 * j + 1] = x.y; av[4 * j + 2] = x.z; av[4 * j + 3] = x.w;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j) av[j] = 0u;
                }
                tmem_st16(lane_base + rb * 32 + h * 16, av);
                tmem_st_wait();
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&afull_bar[rb]);
            }
            // ---- per-list state
            float tk[EV_SEEDK];                               // seed mode: largest half-tile-pair maxima, descending
            float pend = -INFINITY;                           // seed mode: maximum of the pair's first half tile
            float tau = INFINITY;
            bool give_up = false;
            uint32_t *rec_base = nullptr, *rec_p = nullptr, *rec_limit = nullptr;
            if (SEED) {
#pragma unroll
                for (int s = 0; s < EV_SEEDK; ++s) tk[s] = -INFINITY;
            } else {
                // A row with more train items than the seed table covers (a user with hundreds of them) is not
                // pre-filtered at all: it goes to the exact engine.
                give_up = !row_ok;
                if (row_ok && a.mask_rowptr != nullptr) give_up = (a.mask_rowptr[row + 1] - a.mask_rowptr[row]) >= EV_SEED_TAB;
                // Any starting threshold is safe: "every unlisted column scored <= the threshold in force when
                // it was seen" holds either way, and a row left with too few candidates fails certification.
                if (!give_up) tau = a.tau0 ? a.tau0[row] : -INFINITY;
                rec_base = a.recs + ((static_cast<size_t>(row_ok ? row : 0) * a.n_slices + sl) * EV_CH + h) * (EV_GCAP * EV_REC_WORDS);
                rec_p = rec_base;
                rec_limit = rec_base + EV_GLIMIT * EV_REC_WORDS;
            }
            // One visit = one half tile.  The loop is unrolled over the row block's two accumulator slots (tile counts
            // are even), so barrier / TMEM addresses are loop constants and the phase parity flips once per pair.
            auto visit = [&](const uint32_t bar_full, const uint32_t bar_empty, const uint32_t taddr, const int i, const uint32_t tcol) {
#if CHAOREC_EV_SPIN & 2
                if (!mbar_spin_a(bar_full, spar, 64))
#else
                if (!mbar_try_wait_a(bar_full, spar))
#endif
                    ev_wait_slow_a(bar_full, spar, (0x500u << 20) | (static_cast<uint32_t>(i) & 0xFFFFFu));
                tc_fence_after();
                uint32_t v[32];
                __syncwarp();
#if CHAOREC_EV_DEBUG == 3
                (void)taddr;                                  // pipeline floor: no accumulator reads at all
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
#else
                tmem_ld32(taddr, v);
                tmem_ld_wait();
#endif
                // the scores are in registers: hand the slot back BEFORE filtering
                tc_fence_before();
                __syncwarp();
                asm volatile(
                    "{\n\t"
                    ".reg .pred p;\n\t"
                    "setp.ne.u32 p, %1, 0;\n\t"
                    "@p mbarrier.arrive.shared::cta.b64 _, [%0];\n\t"
                    "}" ::"r"(bar_empty), "r"(lane0)
                    : "memory");
                // (zero padding beyond Nb exists in the last column tiles only; those columns score exactly 0 and, if
                // a row lists them, the re-score kernel drops them like any other invalid column)
#if CHAOREC_EV_DEBUG == 2 || CHAOREC_EV_DEBUG == 3
                (void)tcol;                                   // accumulator reads, no filter
                return;
#endif
                float m8[4];
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) m8[gq] = max8(v + 8 * gq);
                const float top = fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3]));
                if (SEED) {
                    // one insertion per PAIR of half tiles (the k-th largest pair maximum is still a lower bound of
                    // the k-th largest sample score): halves the cost of the insertion network
                    if ((i & 1) == 0) {
                        pend = top;
                    } else {
                        float x = fmaxf(top, pend);
#pragma unroll
                        for (int s = 0; s < EV_SEEDK; ++s) {
                            const float hi = fmaxf(tk[s], x);
                            x = fminf(tk[s], x);
                            tk[s] = hi;
                        }
                    }
                } else if (__any_sync(0xffffffffu, top > tau) && CHAOREC_EV_DEBUG != 1) {   // (debug 1: maxima and test only)
                    // Some row of the warp has a candidate in this half tile.  No narrowing down to columns here (that
                    // cost ~45 instructions per hot group for the whole warp, with one lane doing useful work): every
                    // lane stores the 8-column groups whose maximum beats its threshold as they are, under a predicate.
                    // (one group is hot in a typical hot visit: the other three cost a vote and a branch each)
#pragma unroll
                    for (int gq = 0; gq < 4; ++gq)
                        if (__any_sync(0xffffffffu, m8[gq] > tau)) store_group_if_above(rec_p, m8[gq], tau, v + 8 * gq, tcol + 8 * gq);
                    // A list that gets within a visit of its capacity means the starting threshold was far too low for
                    // this row: the row is handed to the exact engine.  (Lists are far shorter than 2 GB: the low
                    // words of the two pointers decide.)
                    if (static_cast<int32_t>(static_cast<uint32_t>(reinterpret_cast<uintptr_t>(rec_p)) -
                                             static_cast<uint32_t>(reinterpret_cast<uintptr_t>(rec_limit))) > 0) {
                        give_up = true;
                        tau = INFINITY;
                    }
                }
            };
            {
                uint32_t tcol = (SEED ? static_cast<uint32_t>(a.seed_off) : static_cast<uint32_t>(t_lo)) * EV_BN + h * 32;
                const uint32_t dcol = (SEED ? static_cast<uint32_t>(a.seed_stride) : 1u) * EV_BN;
                for (int i = 0; i < n_t; i += 2) {
                    visit(bar_full0, bar_empty0, taddr0, i, tcol);
                    visit(bar_full0 + 8, bar_empty0 + 8, taddr0 + EV_BN, i + 1, tcol + dcol);
                    tcol += 2 * dcol;
                    spar ^= 1u;
                }
            }
            // ---- end of unit
            if (SEED) {
                // Both threads of a row publish their sorted half-tile maxima; the k-th largest of the union is a
                // lower bound of the row's k-th largest sample score.  k depends on how many of the row's top
                // columns may turn out to be train items.
                float *mine = a.seed_scratch + ((static_cast<size_t>(blockIdx.x) * EV_BM + row_local) * EV_CH + h) * EV_SEEDK;
#pragma unroll
                for (int s = 0; s < EV_SEEDK; ++s) mine[s] = tk[s];
                __threadfence_block();
                asm volatile("bar.sync 1, %0;" ::"n"(EV_EPI_WARPS * 32) : "memory");     // epilogue warps only
                if (h == 0 && row_ok) {
                    int nmask = 0;
                    if (a.mask_rowptr != nullptr) {
                        const int64_t len = a.mask_rowptr[row + 1] - a.mask_rowptr[row];
                        nmask = len < EV_SEED_TAB - 1 ? static_cast<int>(len) : EV_SEED_TAB - 1;
                    }
                    const int kk = a.seed_k[nmask];
                    const float *other = mine + EV_SEEDK;
                    int ia = 0, ib = 0;
                    float vsel = -INFINITY;
                    for (int s = 0; s < kk; ++s) {            // two-pointer merge of two descending lists
                        const float xa = ia < EV_SEEDK ? mine[ia] : -INFINITY;
                        const float xb = ib < EV_SEEDK ? other[ib] : -INFINITY;
                        if (xa >= xb) { vsel = xa; ++ia; } else { vsel = xb; ++ib; }
                    }
                    a.tau0[row] = vsel;
                }
                asm volatile("bar.sync 1, %0;" ::"n"(EV_EPI_WARPS * 32) : "memory");     // scratch may be overwritten now
            } else if (row_ok) {
                const size_t o = (static_cast<size_t>(row) * a.n_slices + sl) * EV_CH + h;
                a.cnt_out[o] = give_up ? 0 : static_cast<int32_t>((rec_p - rec_base) / EV_REC_WORDS);
                a.tau_out[o] = give_up ? INFINITY : tau;
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == EV_MMA_WARP0) tmem_dealloc(tmem_base, 512);
}

constexpr size_t ev_smem_bytes() {
    return static_cast<size_t>(EV_CTRL_OFF) + sizeof(EvCtrl) + 64;
}

// ---- fp32 -> scaled fp16 operand copies -----------------------------------------------------------------
// x' = x * 2^-e is exact; fp16(x') has relative error <= 2^-11 in the normal range and absolute error <= 2^-25
// below 2^-14.  A rows get their own exponent (a positive per-row factor does not change the row's ranking);
// the item table shares one exponent.  16 lanes per row, float4 each: rows of up to 64 elements.
__global__ void absmax_kernel(const float *__restrict__ x, int64_t n, unsigned int *__restrict__ out_bits) {
    float m = 0.0f;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x)
        m = fmaxf(m, fabsf(x[i]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) atomicMax(out_bits, __float_as_uint(m));   // non-negative floats order as integers
}

__device__ __forceinline__ int scale_exponent(float amax) {
    // e with amax * 2^-e in [1, 2) (0 for an all-zero input); clamped so that 2^-e and 2^e are normal floats
    if (!(amax > 0.0f)) return 0;
    int e = static_cast<int>((__float_as_uint(amax) >> 23) & 0xFFu) - 127;
    if (e < -100) e = -100;
    if (e > 100) e = 100;
    return e;
}
__device__ __forceinline__ float pow2f(int e) { return __uint_as_float(static_cast<uint32_t>(127 + e) << 23); }

// per_row != 0: each row uses its own exponent; else the exponent of *shared_absmax_bits.
// stats (nullable, shared mode): [0] = 2^e as a float, [1] = max scaled row norm (upper bound) via atomicMax.
__global__ void to_f16_scaled_kernel(const float *__restrict__ x, int64_t n_rows, int D, int per_row,
                                     const unsigned int *__restrict__ shared_absmax_bits, __half *__restrict__ out,
                                     float *__restrict__ row_norm_scaled, float *__restrict__ row_scale,
                                     float *__restrict__ stats) {
    const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 4;
    const int sub = threadIdx.x & 15;
    const bool ok = row < n_rows;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ok && sub * 4 < D) v = *reinterpret_cast<const float4 *>(x + row * D + sub * 4);
    float amax = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
    float ss = v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
        amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    const int e = per_row ? scale_exponent(amax) : scale_exponent(__uint_as_float(*shared_absmax_bits));
    const float sc = pow2f(-e);
    if (ok) {
        const __half2 lo = __floats2half2_rn(v.x * sc, v.y * sc), hi = __floats2half2_rn(v.z * sc, v.w * sc);
        uint2 pk;
        pk.x = *reinterpret_cast<const uint32_t *>(&lo);
        pk.y = *reinterpret_cast<const uint32_t *>(&hi);
        *reinterpret_cast<uint2 *>(out + row * EV_KPAD + sub * 4) = pk;
    }
    // fp32 summation error <= D * 2^-24 relative: rounded up, the norm feeds an upper bound
    const float nrm = sqrtf(ss) * sc * (1.0f + 2.0e-7f * static_cast<float>(D));
    if (ok && sub == 0) {
        if (row_norm_scaled) row_norm_scaled[row] = nrm;
        if (row_scale) row_scale[row] = pow2f(e);
        if (stats) {
            atomicMax(reinterpret_cast<unsigned int *>(stats + 1), __float_as_uint(nrm));
            if (row == 0) stats[0] = pow2f(e);
        }
    }
}

// ---- list post-processing + exact re-score + certification: one warp per row ------------------------------
struct RescoreArgs {
    const float *A;            // fp32 operands (canonical arithmetic)
    const float *B;
    int64_t Ma, Nb;
    int D, K, n_slices, kp;
    const uint32_t *recs;      // [Ma][n_slices][EV_GCAP] group records (n_slices = column slices x threads per row)
    const int32_t *cnt;        // [Ma][n_slices] records per list
    const float *tau;          // [Ma][n_slices]
    const float *a_norm;       // [Ma] scaled row norms (upper bounds)
    const float *a_scale;      // [Ma] 2^ea
    const float *b_stats;      // [0] 2^eb, [1] max scaled item norm (upper bound)
    float gamma;               // |approx - canonical| <= gamma * |a'| * |b'|max in scaled units
    const int64_t *mask_rowptr;
    const int32_t *mask_col;
    float mask_value;
    int exclude_self;
    int64_t self_off;
    int64_t id_offset;
    int64_t *out_idx;
    float *out_score;
    int32_t *fallback_rows;    // rows whose exact top-K could not be certified
    uint64_t *fallback_seeds;  // their exact K-th best key so far (0 if fewer than K entries)
    unsigned int *fallback_count;
    const unsigned int *abort_flag;   // != 0: the pre-filter pipeline timed out -> every row goes to the exact engine
};

constexpr int kRescoreWarps = 4;

__global__ void __launch_bounds__(kRescoreWarps * 32, 12) rescore_lists_kernel(const RescoreArgs a, const int keys_per_warp) {
    extern __shared__ __align__(16) unsigned char rs_smem[];   // [warps][keys_per_warp] keys, [warps][256] histogram, [warps][EV_SEED_TAB] train items, [warps] key counts
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int64_t row = static_cast<int64_t>(blockIdx.x) * kRescoreWarps + wib;
    if (row >= a.Ma) return;
    uint64_t *keys = reinterpret_cast<uint64_t *>(rs_smem) + static_cast<size_t>(wib) * keys_per_warp;
    uint32_t *hist = reinterpret_cast<uint32_t *>(reinterpret_cast<uint64_t *>(rs_smem) + static_cast<size_t>(kRescoreWarps) * keys_per_warp) + wib * 256;
    int32_t *mask_s = reinterpret_cast<int32_t *>(hist - wib * 256 + kRescoreWarps * 256) + wib * EV_SEED_TAB;   // the row's train items
    int *n_keys = reinterpret_cast<int *>(mask_s - wib * EV_SEED_TAB + kRescoreWarps * EV_SEED_TAB) + wib;
    // ---- gather the row's records (all column slices and halves): the scores above the list's threshold become keys.
    // A lane takes whole records (two 16-byte loads + the column word, all independent), so a typical row -- ~50
    // records per list -- is two rounds of loads per list; the order of `keys` is irrelevant (they are ranked below).
    float tmain = -INFINITY;        // every valid column outside `keys` has approximate score <= tmain
    bool bad = *a.abort_flag != 0u;
    if (lane == 0) *n_keys = 0;
    __syncwarp();
    for (int s = 0; s < a.n_slices; ++s) {
        const size_t o = static_cast<size_t>(row) * a.n_slices + s;
        const int c = a.cnt[o];
        const float t = a.tau[o];
        if (!(t < INFINITY)) bad = true;
        tmain = fmaxf(tmain, t);
        const uint32_t *src = a.recs + o * (EV_GCAP * EV_REC_WORDS);
#pragma unroll 2
        for (int r = lane; r < c; r += 32) {
            const uint32_t *rec = src + r * EV_REC_WORDS;
            const uint4 x0 = __ldcs(reinterpret_cast<const uint4 *>(rec));
            const uint4 x1 = __ldcs(reinterpret_cast<const uint4 *>(rec) + 1);
            const uint32_t col0 = __ldcs(rec + 8);
            const uint32_t sc[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
            int hits = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) hits += (__uint_as_float(sc[j]) > t) ? 1 : 0;
            int pos = atomicAdd(n_keys, hits);          // one reservation per record (it has one hit, rarely two)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (__uint_as_float(sc[j]) > t) {
                    if (pos < keys_per_warp) keys[pos] = make_key(__uint_as_float(sc[j]), col0 + j);
                    ++pos;
                }
            }
        }
    }
    __syncwarp();
    int n = *n_keys;
    if (n > keys_per_warp) { bad = true; n = keys_per_warp; }
    int64_t m_lo = 0;
    int nm_all = 0;
    if (a.mask_rowptr != nullptr) {
        m_lo = a.mask_rowptr[row];
        nm_all = static_cast<int>(a.mask_rowptr[row + 1] - m_lo);
    }
    // (rows with EV_SEED_TAB or more train items were never pre-filtered: they are `bad` already)
    if (nm_all >= EV_SEED_TAB) bad = true;
    uint64_t kth = 0ull;
    bool ok = false;
    if (!bad) {
        // the row's sorted train items: one coalesced read, then every look-up is a shared-memory binary search
        for (int t = lane; t < nm_all; t += 32) mask_s[t] = a.mask_col[m_lo + t];
        __syncwarp();
        // ---- cut to the r_keep best by approximate score (no sort), then drop invalid columns
        const int r_keep = a.kp + (a.exclude_self ? 1 : 0) + (nm_all < EV_MASK_CAP ? nm_all : EV_MASK_CAP);
        uint32_t prefix = 0u;
        if (n > 64 && n > r_keep + 1) {
            prefix = select_prefix([&](int i) { return keys[i]; }, n, r_keep, lane, hist);
            tmain = fmaxf(tmain, score_of_bits(prefix));     // everything cut scored below the prefix
        }
        const int64_t self_col = a.exclude_self ? row + a.self_off : -1;
        int m = 0;                                            // valid keys, compacted to keys[0..m)
        for (int base = 0; base < n; base += 32) {
            const int i = base + lane;
            const uint64_t key = (i < n) ? keys[i] : 0ull;
            bool v = key != 0ull && static_cast<uint32_t>(key >> 32) >= prefix;
            if (v) {
                const int64_t col = key_index(key);
                if (col == self_col || col >= a.Nb) v = false;
                int lo = 0, hi = nm_all;                      // binary search in the row's sorted train items
                while (lo < hi) {
                    const int mid = (lo + hi) >> 1;
                    const int32_t mc = mask_s[mid];
                    if (mc < col) lo = mid + 1; else hi = mid;
                }
                if (lo < nm_all && mask_s[lo] == col) v = false;
            }
            const unsigned bm = __ballot_sync(0xffffffffu, v);
            __syncwarp();
            if (v) keys[m + __popc(bm & ((1u << lane) - 1u))] = key;
            m += __popc(bm);
            __syncwarp();
        }
        int P = 64;
        while (P < m) P <<= 1;
        if (P > keys_per_warp) { P = 64; m = 0; }             // (cannot happen: m <= n <= keys_per_warp, a power of two)
        for (int t = m + lane; t < P; t += 32) keys[t] = 0ull;
        __syncwarp();
        warp_sort_desc_fast(keys, P, lane);
        int n_cand = m;
        if (m > a.kp) {                                       // only the kp best valid candidates are re-scored
            tmain = fmaxf(tmain, key_score(keys[a.kp]));
            n_cand = a.kp;
        }
        // ---- exact canonical scores of the candidates
        const float *ar = a.A + row * a.D;
        uint64_t mine[EV_KP_MAX / 32];
#pragma unroll
        for (int r = 0; r < EV_KP_MAX / 32; ++r) {
            const int t = lane + r * 32;
            mine[r] = (t < n_cand) ? keys[t] : 0ull;
        }
        __syncwarp();
        int nm = nm_all < a.K ? nm_all : a.K;
        int P2 = 64;
        while (P2 < n_cand + nm) P2 <<= 1;
#pragma unroll 1
        for (int r = 0; r < EV_KP_MAX / 32; ++r) {
            const int t = lane + r * 32;
            if (t >= n_cand) continue;
            const uint32_t col = key_index(mine[r]);
            const float *br = a.B + static_cast<int64_t>(col) * a.D;
            float acc = 0.0f;
            int k = 0;
            // canonical chain: k ascending, one fmaf per term
            for (; k + 32 <= a.D; k += 32) {
                float4 y[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) y[u] = __ldg(reinterpret_cast<const float4 *>(br + k) + u);
#pragma unroll
                for (int u = 0; u < 8; ++u) {
                    const float4 x = *reinterpret_cast<const float4 *>(ar + k + 4 * u);
                    acc = fmaf(x.x, y[u].x, acc);
                    acc = fmaf(x.y, y[u].y, acc);
                    acc = fmaf(x.z, y[u].z, acc);
                    acc = fmaf(x.w, y[u].w, acc);
                }
            }
            for (; k < a.D; k += 4) {
                const float4 x = *reinterpret_cast<const float4 *>(ar + k);
                const float4 y = __ldg(reinterpret_cast<const float4 *>(br + k));
                acc = fmaf(x.x, y.x, acc);
                acc = fmaf(x.y, y.y, acc);
                acc = fmaf(x.z, y.z, acc);
                acc = fmaf(x.w, y.w, acc);
            }
            keys[t] = make_key(acc, col);
        }
        // Common case: at most 32 candidates whose K-th best exact score beats mask_value -- no train item can enter the
        // top K, so the 32 exact keys are ranked in registers and nothing is merged.
        bool fast = false;
        if (n_cand <= 32 && a.K <= n_cand) {
            __syncwarp();
            uint64_t mykey = lane < n_cand ? keys[lane] : 0ull;
            warp_sort32_desc_regs(mykey, lane);
            const uint64_t kth32 = __shfl_sync(0xffffffffu, mykey, a.K - 1);
            fast = nm == 0 || kth32 > make_key(a.mask_value, 0u);       // (every train item's key is <= the key of column 0)
            if (fast) {
                __syncwarp();
                keys[lane] = mykey;
            }
        }
        __syncwarp();
        // the first K train items re-enter with mask_value (Model/LightGCN.py:150-152)
        for (int t = n_cand + lane; t < P2 && !fast; t += 32) {
            uint64_t key = 0ull;
            if (t < n_cand + nm) {
                const int32_t mc = mask_s[t - n_cand];
                if (mc >= 0 && mc < a.Nb) key = make_key(a.mask_value, static_cast<uint32_t>(mc));
            }
            keys[t] = key;
        }
        __syncwarp();
        if (!fast) warp_sort_desc_fast(keys, P2, lane);
        // ---- certification: a valid column outside the candidate set has approximate (scaled) score <= tmain,
        // hence canonical score <= (tmain + eps') * 2^(ea+eb); the exact top-K is final iff its K-th score is
        // strictly above that.  (The power-of-two factors are exact; eps' is rounded up.)
        kth = keys[a.K - 1];
        const float eps = a.gamma * a.a_norm[row] * a.b_stats[1];
        const float bound = (tmain + eps * 1.0001f) * (a.a_scale[row] * a.b_stats[0]);
        const bool complete = (tmain == -INFINITY);
        ok = kth != 0ull && (complete || key_score(kth) > bound);
    }
    if (ok) {
        for (int t = lane; t < a.K; t += 32) {
            const uint64_t key = keys[t];
            a.out_idx[row * a.K + t] = static_cast<int64_t>(key_index(key)) + a.id_offset;
            if (a.out_score) a.out_score[row * a.K + t] = key_score(key);
        }
    } else if (lane == 0) {
        const unsigned slot = atomicAdd(a.fallback_count, 1u);
        a.fallback_rows[slot] = static_cast<int32_t>(row);
        a.fallback_seeds[slot] = kth;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn ev_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
        else
            cudaGetLastError();
    }
    return fn;
}

bool make_map_f16(CUtensorMap *map, const void *base, int64_t rows) {
    EncodeTiledFn fn = ev_encode_fn();
    if (!fn) return false;
    const cuuint64_t gdim[2] = {static_cast<cuuint64_t>(EV_KPAD), static_cast<cuuint64_t>(rows)};
    const cuuint64_t gstride[1] = {static_cast<cuuint64_t>(EV_KPAD) * 2};
    const cuuint32_t box[2] = {static_cast<cuuint32_t>(EV_KPAD), static_cast<cuuint32_t>(EV_BN)};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(base), gdim, gstride, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t al256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

// Smallest k with P(Poisson(x) >= k) <= p.
int poisson_rank(double x, double p) {
    double term = exp(-x), cdf = term;      // P(Poisson <= 0)
    int k = 1;
    while (1.0 - cdf > p && k < 64) { term *= x / k; cdf += term; ++k; }
    return k;
}

}  // namespace

// Column slices only buy load balance: all slices of a row share the row's starting threshold, so a slice's list
// holds its share of the row's candidates.  Estimated time ~ rounds(units / SMs) / slices, plus 2 % per extra
// slice for the additional unit start-ups and list bookkeeping.
int tc_eval_choose_slices(int64_t Ma, int64_t Nb) {
    const int row_tiles = static_cast<int>((Ma + EV_BM - 1) / EV_BM);
    const int col_tiles = static_cast<int>((Nb + EV_BN - 1) / EV_BN);
    const int sms = sm_count();
    int slices = 1;
    double best = 1e30;
    for (int s = 1; s <= 4; ++s) {
        if (s > 1 && col_tiles / s < 64) break;
        const int rounds = (row_tiles * s + sms - 1) / sms;
        const double cost = static_cast<double>(rounds) / s * (1.0 + 0.02 * (s - 1));
        if (cost < best - 1e-9) { best = cost; slices = s; }
    }
    return slices;
}

size_t tc_eval_workspace_bytes(int64_t Ma, int64_t Nb, int D) {
    const int slices = tc_eval_choose_slices(Ma, Nb);
    size_t n = 256;                                                    // counters
    n += al256(static_cast<size_t>(Ma) * EV_KPAD * 2);                  // A fp16
    n += al256(static_cast<size_t>(Nb) * EV_KPAD * 2);                  // B fp16
    n += 2 * al256(static_cast<size_t>(Ma) * 4) + 256;                  // A norms and scales, B stats
    n += al256(static_cast<size_t>(Ma) * 4);                            // starting thresholds
    n += 2048 + al256(static_cast<size_t>(Ma) * slices * EV_CH * EV_GCAP * EV_REC_BYTES);   // group records
    n += 2 * al256(static_cast<size_t>(Ma) * slices * EV_CH * 4);       // list lengths, final thresholds
    n += al256(static_cast<size_t>(sm_count()) * EV_BM * EV_CH * EV_SEEDK * 4);   // seed pass: half-tile maxima of a unit
    n += al256(static_cast<size_t>(Ma) * 4);                            // fallback row list
    n += al256(static_cast<size_t>(Ma) * 8);                            // ... and their seed keys
    n += rank_listed_workspace_bytes(D);                                // exact engine for uncertified rows
    return n;
}

bool tc_eval_supported(int64_t Ma, int64_t Nb, int D, int K) {
    // small item tables stay on the exact engine (nothing to win, and K' must be well below Nb)
    return D % 4 == 0 && D >= 4 && D <= EV_KPAD && K >= 1 && K <= 64 && Nb >= 2048 && Ma >= 1 && Ma < 0x7FFFFFF0LL &&
           Nb < 0x7FFFFFC0LL && ev_encode_fn() != nullptr;
}

namespace {
// The one place that knows how the workspace is laid out (the launcher and the fallback-row query both use it).
struct EvalWs {
    int slices;
    unsigned int *counters;
    __half *a16, *b16;
    float *a_norm, *a_scale, *b_stats, *tau0;
    uint32_t *recs;
    int32_t *cnt;
    float *tau, *seed_scratch;
    int32_t *fb_rows;
    uint64_t *fb_seeds;
    void *rank_ws;
};
EvalWs carve_eval_ws(void *workspace, int64_t Ma, int64_t Nb) {
    EvalWs w{};
    const int slices = tc_eval_choose_slices(Ma, Nb);
    w.slices = slices;
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    w.counters = reinterpret_cast<unsigned int *>(ws); ws += 256;
    w.a16 = reinterpret_cast<__half *>(ws); ws += al256(static_cast<size_t>(Ma) * EV_KPAD * 2);
    w.b16 = reinterpret_cast<__half *>(ws); ws += al256(static_cast<size_t>(Nb) * EV_KPAD * 2);
    w.a_norm = reinterpret_cast<float *>(ws); ws += al256(static_cast<size_t>(Ma) * 4);
    w.a_scale = reinterpret_cast<float *>(ws); ws += al256(static_cast<size_t>(Ma) * 4);
    w.b_stats = reinterpret_cast<float *>(ws); ws += 256;
    w.tau0 = reinterpret_cast<float *>(ws); ws += al256(static_cast<size_t>(Ma) * 4);
    ws = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(ws) + 2047) & ~static_cast<uintptr_t>(2047));
    w.recs = reinterpret_cast<uint32_t *>(ws); ws += al256(static_cast<size_t>(Ma) * slices * EV_CH * EV_GCAP * EV_REC_BYTES);
    w.cnt = reinterpret_cast<int32_t *>(ws); ws += al256(static_cast<size_t>(Ma) * slices * EV_CH * 4);
    w.tau = reinterpret_cast<float *>(ws); ws += al256(static_cast<size_t>(Ma) * slices * EV_CH * 4);
    w.seed_scratch = reinterpret_cast<float *>(ws); ws += al256(static_cast<size_t>(sm_count()) * EV_BM * EV_CH * EV_SEEDK * 4);
    w.fb_rows = reinterpret_cast<int32_t *>(ws); ws += al256(static_cast<size_t>(Ma) * 4);
    w.fb_seeds = reinterpret_cast<uint64_t *>(ws); ws += al256(static_cast<size_t>(Ma) * 8);
    w.rank_ws = ws;
    return w;
}
}  // namespace

// Device pointer of the list of rows the last tc_eval_launch on this workspace handed to the exact engine (their
// number is the uint32 at byte offset 64 of the workspace).
const int32_t *tc_eval_fallback_rows(void *workspace, int64_t Ma, int64_t Nb) { return carve_eval_ws(workspace, Ma, Nb).fb_rows; }

// A, B: fp32 operands (same pointer allowed).  Computes the exact canonical top-K of every row of A.
int tc_eval_launch(const EngineArgs &e, void *workspace, size_t workspace_bytes, cudaStream_t stm) {
    if (workspace_bytes < tc_eval_workspace_bytes(e.Ma, e.Nb, e.d)) return CHAOREC_E_WORKSPACE;
    const EvalWs w = carve_eval_ws(workspace, e.Ma, e.Nb);
    const int slices = w.slices;
    unsigned int *counters = w.counters;
    __half *a16 = w.a16, *b16 = w.b16;
    float *a_norm = w.a_norm, *a_scale = w.a_scale, *b_stats = w.b_stats, *tau0 = w.tau0, *tau = w.tau, *seed_scratch = w.seed_scratch;
    uint32_t *recs = w.recs;
    int32_t *cnt = w.cnt, *fb_rows = w.fb_rows;
    uint64_t *fb_seeds = w.fb_seeds;
    void *rank_ws = w.rank_ws;
    const size_t rank_ws_bytes = workspace_bytes - static_cast<size_t>(static_cast<unsigned char *>(w.rank_ws) - static_cast<unsigned char *>(workspace));

    // counters[16] = fallback count (read by the host shim), counters[32] = max |B| bits; b_stats zeroed for atomicMax
    if (cudaMemsetAsync(counters, 0, 256, stm) != cudaSuccess || cudaMemsetAsync(b_stats, 0, 256, stm) != cudaSuccess) {
        cudaGetLastError();
        return CHAOREC_E_LAUNCH;
    }
    {
        const int64_t nb_el = e.Nb * e.d;
        int64_t blocks = (nb_el + 256 * 8 - 1) / (256 * 8);
        if (blocks > 4096) blocks = 4096;
        absmax_kernel<<<static_cast<unsigned>(blocks), 256, 0, stm>>>(e.B, nb_el, counters + 32);
        to_f16_scaled_kernel<<<static_cast<unsigned>((e.Ma * 16 + 255) / 256), 256, 0, stm>>>(e.A, e.Ma, e.d, 1, nullptr, a16, a_norm,
                                                                                              a_scale, nullptr);
        to_f16_scaled_kernel<<<static_cast<unsigned>((e.Nb * 16 + 255) / 256), 256, 0, stm>>>(e.B, e.Nb, e.d, 0, counters + 32, b16,
                                                                                              nullptr, nullptr, b_stats);
    }
    CUtensorMap map_b;
    if (!make_map_f16(&map_b, b16, e.Nb)) return CHAOREC_E_LAUNCH;

    int kp = e.K <= 20 ? 32 : (e.K <= 50 ? 64 : 128);
    if (e.kp_hint == 32 || e.kp_hint == 64 || e.kp_hint == 128) kp = e.kp_hint;
    if (const char *env = getenv("CHAOREC_TC_KP")) {
        const int v = atoi(env);
        if (v == 32 || v == 64 || v == 128) kp = v;
    }
    if (kp < e.K + 8) kp = (e.K + 8 <= 64) ? 64 : 128;

    EvalArgs t{};
    t.Ma = e.Ma; t.Nb = e.Nb;
    t.n_row_tiles = static_cast<int>((e.Ma + EV_BM - 1) / EV_BM);
    // (tile counts are even everywhere: the epilogue's tile loop is unrolled over the row block's two accumulator
    // slots; a padding tile beyond Nb scores 0 everywhere and the re-score kernel drops its columns)
    const int real_col_tiles = static_cast<int>((e.Nb + EV_BN - 1) / EV_BN);
    t.n_col_tiles = (real_col_tiles + 1) & ~1;
    t.n_slices = slices;
    t.tiles_per_slice = (((t.n_col_tiles + slices - 1) / slices) + 1) & ~1;
    t.a16 = a16;
    t.mask_rowptr = e.mask_rowptr;
    t.exclude_self = e.exclude_self;
    t.kp = kp;
    t.recs = recs; t.cnt_out = cnt; t.tau_out = tau; t.tau0 = nullptr;
    t.dbg = counters + 36;
    t.seed_scratch = seed_scratch;

    const size_t smem = ev_smem_bytes();
    cudaError_t ce = cudaFuncSetAttribute(tc_eval_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce == cudaSuccess)
        ce = cudaFuncSetAttribute(tc_eval_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (ce != cudaSuccess) { cudaGetLastError(); return CHAOREC_E_LAUNCH; }

    // ---- seed pass: starting thresholds from an evenly spaced sample of column tiles ------------------------
    // The k-th largest sample score sits, in expectation, at rank k / f of the full row (f = sampled fraction);
    // k is the smallest rank that leaves >= need full-row columns above the threshold with probability
    // 1 - 1e-3 (a Poisson(need * f) variable reaching k), need = K' + a typical train-item list.  Rows left
    // short fail certification and go to the exact engine.  The sample is sized so that k <= 16 tile maxima.
    int seed_tiles = 512;
    if (const char *env = getenv("CHAOREC_TC_SEED_TILES")) seed_tiles = atoi(env);
    // columns a row needs above its threshold: K, the columns inside the certification window and the slack the
    // approximate order needs (6), plus every train item of the row (a trained model ranks them at the top)
    const double need0 = e.K + 6.0 + (e.exclude_self ? 1.0 : 0.0);
    {
        const int cap = static_cast<int>(8.5 / (need0 + 16.0) * real_col_tiles);   // (need0 + 16) * f <= 8.5 -> k <= 20
        if (seed_tiles > cap) seed_tiles = cap;
        if (seed_tiles > real_col_tiles / 2) seed_tiles = real_col_tiles / 2;
        seed_tiles &= ~1;
    }
    if (seed_tiles >= 2) {
        EvalArgs p = t;
        p.n_slices = 1;
        p.tiles_per_slice = t.n_col_tiles;
        p.n_seed_tiles = seed_tiles;
        p.seed_stride = real_col_tiles / seed_tiles;
        p.seed_off = p.seed_stride / 2;
        const double f = static_cast<double>(seed_tiles) * EV_BN / static_cast<double>(e.Nb);
        const int k_max = seed_tiles < EV_SEEDK ? seed_tiles : EV_SEEDK;
        for (int m = 0; m < EV_SEED_TAB; ++m) {
            int k_seed = poisson_rank((need0 + m) * f, 1e-3);
            if (k_seed < 2) k_seed = 2;
            if (k_seed > k_max) k_seed = k_max;      // (rows that would need more are left short and fall back)
            p.seed_k[m] = static_cast<unsigned char>(k_seed);
        }
        p.tau0 = tau0;
        const int pgrid = p.n_row_tiles < sm_count() ? p.n_row_tiles : sm_count();
        tc_eval_kernel<true><<<pgrid, EV_THREADS, smem, stm>>>(map_b, p);
        t.tau0 = tau0;
    }
    {
        const int n_units = t.n_row_tiles * slices;
        const int grid = n_units < sm_count() ? n_units : sm_count();
        tc_eval_kernel<false><<<grid, EV_THREADS, smem, stm>>>(map_b, t);
    }

    RescoreArgs r{};
    r.A = e.A; r.B = e.B; r.Ma = e.Ma; r.Nb = e.Nb; r.D = e.d; r.K = e.K; r.n_slices = slices * EV_CH; r.kp = kp;
    r.recs = recs; r.cnt = cnt; r.tau = tau; r.a_norm = a_norm; r.a_scale = a_scale; r.b_stats = b_stats;
    // |s~' - s'| <= gamma |a'||b'|max in scaled units, against the CANONICAL fp32 score s':
    //   fp16 rounding of both operands      2 * 2^-11 + 2^-22
    //   fp16 subnormal flush (|x'| < 2^-14)  2^-24 * sqrt(64) relative to |a'||b'|max (both >= 1)
    //   tensor-core fp32 accumulation        <= 2 d 2^-23   (truncation per addition assumed at worst)
    //   canonical chain's own rounding       <= d 2^-24
    r.gamma = (9.765625e-4f + 2.4e-7f + 4.8e-7f) + 2.5f * static_cast<float>(e.d) * 1.1920929e-7f;
    r.mask_rowptr = e.mask_rowptr; r.mask_col = e.mask_col; r.mask_value = e.mask_value;
    r.exclude_self = e.exclude_self; r.self_off = e.self_off;
    r.id_offset = e.id_offset; r.out_idx = e.out_idx; r.out_score = e.out_score;
    r.fallback_rows = fb_rows; r.fallback_seeds = fb_seeds; r.fallback_count = counters + 16;
    r.abort_flag = counters + 36;
    {
        // keys of a row: a power of two; a row that lists more (far more than any sane threshold produces) goes to
        // the exact engine
        // Keys per row.  256 keeps 12 CTAs per SM resident (measured: 4.8 ms against 6.5 ms with 384 or 512 at 1M x 200k);
        // a row that lists more goes to the exact engine.  Small item tables get their thresholds from a handful of
        // sample tiles and do produce rows with 300+ keys (431 of 5000 rows at 6000 items): they keep 512.
        const int keys_per_warp = e.Nb >= 32768 ? 256 : 512;
        const size_t rs_smem = static_cast<size_t>(kRescoreWarps) * (keys_per_warp * 8 + 256 * 4 + EV_SEED_TAB * 4 + 4);
        ce = cudaFuncSetAttribute(rescore_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(rs_smem));
        if (ce != cudaSuccess) { cudaGetLastError(); return CHAOREC_E_LAUNCH; }
        rescore_lists_kernel<<<static_cast<unsigned>((e.Ma + kRescoreWarps - 1) / kRescoreWarps), kRescoreWarps * 32, rs_smem, stm>>>(
            r, keys_per_warp);
    }
    int rc = launch_status();
    if (rc != CHAOREC_OK) return rc;
    // rows that could not be certified: exact CUDA-core engine over the listed rows only
    EngineArgs fb = e;
    fb.row_list = fb_rows;
    fb.row_count_dev = counters + 16;
    fb.seed_keys = fb_seeds;
    return rank_launch_engine(fb, rank_ws, rank_ws_bytes, stm);
}

}  // namespace chaorec
