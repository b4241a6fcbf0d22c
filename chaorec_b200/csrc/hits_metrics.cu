// Evaluation tail: top-K ids -> per-user hit bit masks (device), then the five metrics of
// metrics.py:13-57 reduced exactly as utils.gene_metrics does (host, sequential fp64).
#include "common.cuh"

namespace chaorec {
namespace {

// 64 threads (two warps) per val_data row: thread r tests rank r by binary search in the row's
// sorted positives; each warp writes one 32-bit half of the mask, so no zero-fill or atomics.
__global__ void __launch_bounds__(256) topk_hits_kernel(const int64_t *__restrict__ topk_idx,
                                                        const int64_t *__restrict__ row_user,
                                                        const int64_t *__restrict__ pos_rowptr,
                                                        const int64_t *__restrict__ pos_col, int64_t U, int K,
                                                        uint32_t *__restrict__ hit_halves) {
    const int64_t u = static_cast<int64_t>(blockIdx.x) * 4 + (threadIdx.x >> 6);  // row of val_data
    const int r = threadIdx.x & 63;
    bool found = false;
    if (u < U && r < K) {
        const int64_t user = row_user ? row_user[u] : u;
        const int64_t item = topk_idx[user * K + r];
        int64_t lo = pos_rowptr[u], hi = pos_rowptr[u + 1];
        while (lo < hi) {
            const int64_t mid = (lo + hi) >> 1;
            const int64_t v = pos_col[mid];
            if (v == item) { found = true; break; }
            if (v < item) lo = mid + 1; else hi = mid;
        }
    }
    const unsigned bits = __ballot_sync(0xffffffffu, found);
    if ((threadIdx.x & 31) == 0 && u < U) hit_halves[u * 2 + (r >> 5)] = bits;
}

}  // namespace
}  // namespace chaorec

using namespace chaorec;

extern "C" int chaorec_topk_hits(const int64_t *topk_idx, const int64_t *row_user, const int64_t *pos_rowptr,
                                 const int64_t *pos_col, int64_t U, int32_t K, uint64_t *hit_mask, void *stream) {
    if (!topk_idx || !pos_rowptr || !hit_mask || U < 0) return CHAOREC_E_BADARG;
    if (K < 1 || K > 64) return CHAOREC_E_UNSUPPORTED;
    if (U == 0) return CHAOREC_OK;
    const int64_t blocks = (U + 3) / 4;
    if (blocks > 0x7fffffffLL) return CHAOREC_E_UNSUPPORTED;
    topk_hits_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        topk_idx, row_user, pos_rowptr, pos_col, U, K, reinterpret_cast<uint32_t *>(hit_mask));
    return launch_status();
}

// Host-side reduction.  Every operation below is the fp64 operation the python code performs,
// in the same order: `metrics[k][name] += value` per row (utils.py:122-127), `/= num_users` at
// the end (utils.py:132-137).
extern "C" int chaorec_metrics_from_hits(const uint64_t *hit_mask, const int32_t *n_pos_of_row, int64_t n_rows,
                                         const int32_t *ks, int32_t n_k, const double *inv_log, double *out) {
    if (!hit_mask || !n_pos_of_row || !ks || !inv_log || !out || n_rows < 0 || n_k < 0)
        return CHAOREC_E_BADARG;
    for (int t = 0; t < n_k; ++t)
        if (ks[t] < 1 || ks[t] > 64) return CHAOREC_E_UNSUPPORTED;
    for (int t = 0; t < n_k * 5; ++t) out[t] = 0.0;
    // idcg(lim) = sum([inv_log[i] for i in range(lim)]) accumulated left to right (metrics.py:31):
    // the sequential prefix sums are exactly those partial sums
    double idcg_prefix[65];
    idcg_prefix[0] = 0.0;
    int max_k = 0;
    for (int t = 0; t < n_k; ++t) max_k = ks[t] > max_k ? ks[t] : max_k;
    for (int i = 0; i < max_k; ++i) idcg_prefix[i + 1] = idcg_prefix[i] + inv_log[i];
    const uint64_t any_mask = (max_k == 64) ? ~0ull : ((1ull << max_k) - 1ull);
    for (int64_t row = 0; row < n_rows; ++row) {
        const uint64_t mask = hit_mask[row];
        // A row without a hit inside max(k) adds +0 to every running sum (0/k, 0/len, 0/idcg, 0, 0/len):
        // x + 0.0 == x exactly, so skipping it leaves every sum bit-identical.
        if ((mask & any_mask) == 0ull) continue;
        const int np = n_pos_of_row[row];
        for (int t = 0; t < n_k; ++t) {
            const int k = ks[t];
            uint64_t m = (k == 64) ? mask : (mask & ((1ull << k) - 1ull));
            if (m == 0ull) continue;
            const int inter = __builtin_popcountll(m);
            double *o = out + t * 5;
            o[0] += static_cast<double>(inter) / static_cast<double>(k);              // metrics.py:16
            if (np > 0) o[1] += static_cast<double>(inter) / static_cast<double>(np);  // metrics.py:20-23
            if (np > 0) {                                                              // metrics.py:28-39
                const double idcg = idcg_prefix[np < k ? np : k];
                double dcg = 0.0, ap = 0.0;
                int n_hit = 0;
                while (m) {                                                            // hits in rank order
                    const int i = __builtin_ctzll(m);
                    m &= m - 1ull;
                    dcg += inv_log[i];
                    ++n_hit;
                    ap += static_cast<double>(n_hit) / static_cast<double>(i + 1);     // metrics.py:53-56
                }
                o[2] += dcg / idcg;
                o[4] += ap / static_cast<double>(np);                                  // metrics.py:57
            }
            o[3] += 1.0;                                                               // metrics.py:44
        }
    }
    if (n_rows > 0)
        for (int t = 0; t < n_k * 5; ++t) out[t] /= static_cast<double>(n_rows);
    return CHAOREC_OK;
}
