// User-user co-occurrence graph: dualgnn-gen-u-u-matrix.py:20-50 builds
// C[u][v] = |items(u) ∩ items(v)| (u != v) with an O(U^2) Python double loop over sets, then keeps the
// min(200, nnz) largest entries of every row (dualgnn-gen-u-u-matrix.py:85-103).
//
// Here: one CTA per user u.  Counters for all users live in shared memory (two 16-bit counters per word);
// for every item of u the CTA walks the item's user list and bumps the counters (shared-memory atomics), so
// the work is sum_i deg(i)^2 integer increments instead of U^2 set intersections.  The row's top-k is then
// selected without sorting the row: a binary search on the count value finds the cut t* (largest t with
// #(count >= t) >= k), everything above t* is taken, ties at t* are taken in ascending user order (the
// canonical tie-break of this library; torch.topk's order among equal counts is unspecified), and only the
// <= k selected entries are sorted (count descending, user ascending).
#include "common.cuh"

namespace chaorec {
namespace {

constexpr int kCoThreads = 1024;
constexpr int kCoMaxK = 256;

__device__ __forceinline__ unsigned co_count(const unsigned *cnt, int v) { return (cnt[v >> 1] >> ((v & 1) * 16)) & 0xFFFFu; }

__device__ __forceinline__ int block_sum(int x, int *red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    __syncthreads();
    if (lane == 0) red[warp] = x;
    __syncthreads();
    int s = 0;
    if (warp == 0) {
        s = red[lane];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) red[32] = s;
    }
    __syncthreads();
    return red[32];
}

__global__ void __launch_bounds__(kCoThreads) cooccur_topk_kernel(const int64_t *__restrict__ u_rowptr, const int32_t *__restrict__ u_col,
                                                                  const int64_t *__restrict__ i_rowptr, const int32_t *__restrict__ i_col,
                                                                  int U, int k, int64_t row_lo, int64_t row_hi,
                                                                  int32_t *__restrict__ out_idx, int32_t *__restrict__ out_cnt,
                                                                  int32_t *__restrict__ out_n) {
    extern __shared__ unsigned co_smem[];
    unsigned *cnt = co_smem;                               // ceil(U/2) words
    const int words = (U + 1) >> 1;
    uint64_t *sel = reinterpret_cast<uint64_t *>(co_smem + ((words + 1) & ~1));   // kCoMaxK keys
    int *red = reinterpret_cast<int *>(sel + kCoMaxK);     // 33 ints
    int *misc = red + 40;                                  // [0] selected so far
    for (int64_t u = row_lo + blockIdx.x; u < row_hi; u += gridDim.x) {
        for (int w = threadIdx.x; w < words; w += kCoThreads) cnt[w] = 0u;
        if (threadIdx.x < kCoMaxK) sel[threadIdx.x] = 0ull;
        if (threadIdx.x == 0) misc[0] = 0;
        __syncthreads();
        // ---- count: warp per item of u, lanes over the item's users ----
        const int64_t it_lo = u_rowptr[u], it_hi = u_rowptr[u + 1];
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int64_t e = it_lo + warp; e < it_hi; e += kCoThreads / 32) {
            const int64_t item = u_col[e];
            const int64_t lo = i_rowptr[item], hi = i_rowptr[item + 1];
            for (int64_t j = lo + lane; j < hi; j += 32) {
                const int v = i_col[j];
                if (v != u) atomicAdd(&cnt[v >> 1], 1u << ((v & 1) * 16));
            }
        }
        __syncthreads();
        // ---- nnz and max ----
        int my_nz = 0;
        unsigned my_max = 0;
        for (int v = threadIdx.x; v < U; v += kCoThreads) {
            const unsigned c = co_count(cnt, v);
            my_nz += c > 0;
            my_max = max(my_max, c);
        }
        const int nnz = block_sum(my_nz, red);
        const int kk = nnz < k ? nnz : k;
        // block max via the same reduction trick on bit planes would cost more than this: use atomicMax on smem
        if (threadIdx.x == 0) misc[1] = 0;
        __syncthreads();
        if (my_max) atomicMax(&misc[1], static_cast<int>(my_max));
        __syncthreads();
        const int maxc = misc[1];
        if (kk > 0) {
            // ---- largest t with #(count >= t) >= kk ----
            int lo_t = 1, hi_t = maxc;                     // invariant: N(lo_t) >= kk
            while (lo_t < hi_t) {
                const int mid = (lo_t + hi_t + 1) >> 1;
                int c = 0;
                for (int v = threadIdx.x; v < U; v += kCoThreads) c += co_count(cnt, v) >= static_cast<unsigned>(mid);
                const int n_mid = block_sum(c, red);
                if (n_mid >= kk) lo_t = mid; else hi_t = mid - 1;
            }
            const unsigned t = static_cast<unsigned>(lo_t);
            // ---- everything above t (unordered) ----
            for (int v = threadIdx.x; v < U; v += kCoThreads) {
                const unsigned c = co_count(cnt, v);
                if (c > t) {
                    const int slot = atomicAdd(&misc[0], 1);
                    sel[slot] = (static_cast<uint64_t>(c) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(v));
                }
            }
            __syncthreads();
            // ---- ties at t in ascending user order until kk entries are selected ----
            int filled = misc[0];
            for (int base = 0; base < U && filled < kk; base += kCoThreads) {
                const int v = base + threadIdx.x;
                const int flag = (v < U && co_count(cnt, v) == t) ? 1 : 0;
                // block-wide exclusive scan of flag
                int incl = flag;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const int x = __shfl_up_sync(0xffffffffu, incl, o);
                    if (lane >= o) incl += x;
                }
                __syncthreads();
                if (lane == 31) red[warp] = incl;
                __syncthreads();
                if (warp == 0) {
                    int w = red[lane], wi = w;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int x = __shfl_up_sync(0xffffffffu, wi, o);
                        if (lane >= o) wi += x;
                    }
                    red[lane] = wi - w;                    // exclusive prefix of the warp totals
                    if (lane == 31) red[32] = wi;
                }
                __syncthreads();
                const int pos = filled + red[warp] + incl - flag;
                if (flag && pos < kk) sel[pos] = (static_cast<uint64_t>(t) << 32) | (0xFFFFFFFFu - static_cast<unsigned>(v));
                filled += red[32];
                __syncthreads();
            }
            // ---- sort the selected keys (descending) ----
            for (int sz = 2; sz <= kCoMaxK; sz <<= 1) {
                for (int j = sz >> 1; j > 0; j >>= 1) {
                    if (threadIdx.x < kCoMaxK / 2) {
                        const int tt = threadIdx.x;
                        const int i = ((tt & ~(j - 1)) << 1) | (tt & (j - 1));
                        const int l = i | j;
                        const bool desc = ((i & sz) == 0);
                        const uint64_t a = sel[i], b = sel[l];
                        if ((a < b) == desc && a != b) { sel[i] = b; sel[l] = a; }
                    }
                    __syncthreads();
                }
            }
        }
        __syncthreads();
        for (int j = threadIdx.x; j < k; j += kCoThreads) {
            const uint64_t key = (j < kk) ? sel[j] : 0ull;
            out_idx[u * k + j] = (j < kk) ? static_cast<int32_t>(0xFFFFFFFFu - static_cast<unsigned>(key & 0xFFFFFFFFull)) : -1;
            out_cnt[u * k + j] = (j < kk) ? static_cast<int32_t>(key >> 32) : 0;
        }
        if (threadIdx.x == 0) out_n[u] = kk;
        __syncthreads();
    }
}

size_t cooccur_smem_bytes(int64_t U) {
    const size_t words = static_cast<size_t>((U + 1) / 2);
    return ((words + 1) & ~static_cast<size_t>(1)) * 4 + kCoMaxK * 8 + 48 * 4;
}

}  // namespace
}  // namespace chaorec

using namespace chaorec;

extern "C" int64_t chaorec_cooccurrence_max_users(void) { return ((220 * 1024 - kCoMaxK * 8 - 48 * 4 - 8) / 4) * 2; }

extern "C" int chaorec_cooccurrence_topk(const int64_t *user_rowptr, const int32_t *user_items, const int64_t *item_rowptr,
                                         const int32_t *item_users, int64_t U, int64_t I, int32_t k, int64_t row_lo,
                                         int64_t row_hi, int32_t *out_idx, int32_t *out_cnt, int32_t *out_n, void *stream) {
    if (!user_rowptr || !item_rowptr || !out_idx || !out_cnt || !out_n || U <= 0 || I <= 0) return CHAOREC_E_BADARG;
    if (row_lo < 0 || row_hi > U || row_lo > row_hi) return CHAOREC_E_BADARG;
    if (k < 1 || k > kCoMaxK || U > chaorec_cooccurrence_max_users()) return CHAOREC_E_UNSUPPORTED;
    if (row_lo == row_hi) return CHAOREC_OK;
    if (!user_items || !item_users) return CHAOREC_E_BADARG;
    const size_t smem = cooccur_smem_bytes(U);
    if (cudaFuncSetAttribute(cooccur_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) {
        cudaGetLastError();
        return CHAOREC_E_LAUNCH;
    }
    const int64_t rows = row_hi - row_lo;
    const int grid = static_cast<int>(rows < 4LL * sm_count() ? rows : 4LL * sm_count());
    cooccur_topk_kernel<<<grid, kCoThreads, smem, static_cast<cudaStream_t>(stream)>>>(
        user_rowptr, user_items, item_rowptr, item_users, static_cast<int>(U), k, row_lo, row_hi, out_idx, out_cnt, out_n);
    return launch_status();
}
