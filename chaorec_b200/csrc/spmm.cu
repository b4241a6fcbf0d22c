// K1: CSR SpMM with fused degree scaling and layer-mean accumulation (fp32).
//
// Replaces torch.sparse.mm(adj, x) (Model/FREEDOM.py:168,174) and the PyG gather/scatter-add
// of LightGCNConv (Model/LightGCN.py:28-46).  HBM-bound gather of 4*d-byte embedding rows:
//   * a sub-warp of LANES lanes owns one destination row; each lane carries NV float4
//     accumulators, so one 64-d row is one fully coalesced 256-byte request of 16 lanes
//     (two destination rows per warp);
//   * column indices / edge weights are read coalesced once per LANES edges and broadcast with
//     warp shuffles; the gathers are issued four at a time before the FMAs that consume them;
//   * effective edge weight row_scale[r] * val[e] * col_scale[c] is formed in registers, so
//     D^-1/2 A D^-1/2 never exists as an edge-value array (V1-V5 of SURVEY.md §2.3 are all
//     expressible);
//   * the epilogue writes y and/or folds alpha*y into the running layer mean;
//   * hub rows (Zipf item rows with 10^3..10^6 neighbours) are cut into fixed chunks, each
//     reduced by a whole CTA through shared memory into a partial row, then summed in chunk
//     order -> deterministic, no atomics, no serial tail.
#include <cstdlib>
#include "common.cuh"

namespace chaorec {
namespace {

constexpr int kThreads = 256;
constexpr int kHubThreshold = 1024;   // defaults (large graphs); chaorec_spmm_hub_plan scales them down
constexpr int kHubChunk = 4096;
constexpr int kHubMaxChunksPerRow = 64;

// Edges per chunk of a hub row with `deg` edges: at least `hub_chunk`, grown (in steps of 64) so that no
// row has more than kHubMaxChunksPerRow chunks -- the finish kernel then sums at most that many partials.
// The host plan (propagate.py: CSRGraph.plan_hubs) computes the same number.
__host__ __device__ __forceinline__ int64_t hub_chunk_edges(int64_t deg, int hub_chunk) {
    const int64_t per = (deg + kHubMaxChunksPerRow - 1) / kHubMaxChunksPerRow;
    return per > hub_chunk ? (per + 63) / 64 * 64 : hub_chunk;
}

constexpr int kMaxPeers = 8;   // GPUs of one NVSwitch domain

struct SpmmArgs {
    const int64_t *rowptr;
    const int32_t *col;
    const float *val;
    const float *row_scale;
    const float *col_scale;
    const float4 *x;
    float4 *y;
    float4 *acc;
    float alpha;
    int acc_mode;
    const int32_t *hub_rows;
    const int64_t *hub_chunk_ptr;
    int n_hub;
    int hub_threshold;   // rows with more edges than this belong to the hub kernels
    int hub_chunk;       // minimum edges per hub chunk (see hub_chunk_edges)
    float4 *hub_partials;
    const int32_t *row_order;   // nullable: sub-warp i takes row row_order[i] (rows sorted by length)
    int *hub_counters;   // non-null: single-launch mode, the last chunk CTA of a row finishes it
    int64_t n_hub_chunks;
    int64_t M;
    int dq;  // d / 4
    // Peer scatter (multi-GPU item partials, chaorec_spmm_csr_push_f32): row r is not stored to y but into the
    // slot buffer of the GPU that owns it, peer_y[r / peer_block] + ((peer_slot * peer_block + r % peer_block) * d),
    // a plain store over NVLink issued by the epilogue of the row (or of the hub-finish step).  n_peers == 0: off.
    float4 *peer_y[kMaxPeers];
    int n_peers;
    int peer_slot;
    int64_t peer_block;
};

template <int LANES>
__device__ __forceinline__ unsigned group_mask(int lane) {
    constexpr unsigned kOnes = (LANES >= 32) ? 0xffffffffu : ((1u << (LANES & 31)) - 1u);
    return (LANES >= 32) ? kOnes : (kOnes << ((lane / LANES) * LANES));
}

// Accumulate edges [e_lo, e_hi) of one row into acc[NV]; all LANES lanes of the group call it
// with identical bounds.
template <int LANES, int NV, bool WIDE>
__device__ __forceinline__ void gather_edges(const SpmmArgs &a, int64_t e_lo, int64_t e_hi, int gl, unsigned mask,
                                             float4 (&acc)[NV]) {
    for (int64_t e0 = e_lo; e0 < e_hi; e0 += LANES) {
        const int64_t e = e0 + gl;
        int c = 0;
        float w = 0.0f;
        if (e < e_hi) {
            c = __ldcs(a.col + e);                       // streamed once: do not displace gathered rows in L2
            w = a.val ? __ldcs(a.val + e) : 1.0f;
            if (a.col_scale) w *= __ldg(a.col_scale + c);
        }
        const int cnt = (e_hi - e0 < LANES) ? static_cast<int>(e_hi - e0) : LANES;
        int j = 0;
        // eight gathers in flight per sub-warp (short rows -- a user's ~12 items -- finish in two rounds)
        for (; WIDE && j + 8 <= cnt; j += 8) {
            int cj[8];
            float wj[8];
            float4 xv[8][NV];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                cj[t] = __shfl_sync(mask, c, j + t, LANES);
                wj[t] = __shfl_sync(mask, w, j + t, LANES);
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float4 *xr = a.x + static_cast<int64_t>(cj[t]) * a.dq;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int q = gl + v * LANES;
                    xv[t][v] = (q < a.dq) ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int t = 0; t < 8; ++t) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc[v].x = fmaf(wj[t], xv[t][v].x, acc[v].x);
                    acc[v].y = fmaf(wj[t], xv[t][v].y, acc[v].y);
                    acc[v].z = fmaf(wj[t], xv[t][v].z, acc[v].z);
                    acc[v].w = fmaf(wj[t], xv[t][v].w, acc[v].w);
                }
            }
        }
        for (; j + 4 <= cnt; j += 4) {
            int cj[4];
            float wj[4];
            float4 xv[4][NV];
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                cj[t] = __shfl_sync(mask, c, j + t, LANES);
                wj[t] = __shfl_sync(mask, w, j + t, LANES);
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const float4 *xr = a.x + static_cast<int64_t>(cj[t]) * a.dq;
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    const int q = gl + v * LANES;
                    xv[t][v] = (q < a.dq) ? __ldg(xr + q) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int t = 0; t < 4; ++t) {
#pragma unroll
                for (int v = 0; v < NV; ++v) {
                    acc[v].x = fmaf(wj[t], xv[t][v].x, acc[v].x);
                    acc[v].y = fmaf(wj[t], xv[t][v].y, acc[v].y);
                    acc[v].z = fmaf(wj[t], xv[t][v].z, acc[v].z);
                    acc[v].w = fmaf(wj[t], xv[t][v].w, acc[v].w);
                }
            }
        }
        for (; j < cnt; ++j) {
            const int cj = __shfl_sync(mask, c, j, LANES);
            const float wj = __shfl_sync(mask, w, j, LANES);
            const float4 *xr = a.x + static_cast<int64_t>(cj) * a.dq;
#pragma unroll
            for (int v = 0; v < NV; ++v) {
                const int q = gl + v * LANES;
                if (q < a.dq) {
                    const float4 xv = __ldg(xr + q);
                    acc[v].x = fmaf(wj, xv.x, acc[v].x);
                    acc[v].y = fmaf(wj, xv.y, acc[v].y);
                    acc[v].z = fmaf(wj, xv.z, acc[v].z);
                    acc[v].w = fmaf(wj, xv.w, acc[v].w);
                }
            }
        }
    }
}

// y = row_scale * acc; optional store; optional layer-mean update.
template <int LANES, int NV>
__device__ __forceinline__ void row_epilogue(const SpmmArgs &a, int64_t r, int gl, float4 (&acc)[NV], bool nonempty = true) {
    const float rs = a.row_scale ? __ldg(a.row_scale + r) : 1.0f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int q = gl + v * LANES;
        if (q >= a.dq) continue;
        float4 yv = acc[v];
        if (a.row_scale) {
            yv.x *= rs; yv.y *= rs; yv.z *= rs; yv.w *= rs;
        }
        const int64_t o = r * a.dq + q;
        if (a.y) a.y[o] = yv;                            // next layer gathers these rows: normal policy
        if (a.n_peers > 0 && nonempty) {                 // (a row without edges is not sent: the owner's inbox is zero there)
            const uint32_t owner = static_cast<uint32_t>(r) / static_cast<uint32_t>(a.peer_block);   // M < 2^31 in push mode
            const uint32_t b = static_cast<uint32_t>(r) - owner * static_cast<uint32_t>(a.peer_block);
            a.peer_y[owner][(static_cast<int64_t>(a.peer_slot) * a.peer_block + b) * a.dq + q] = yv;
            // The store must have LANDED before this kernel counts as complete: the consumer is released by a flag that
            // a later kernel writes over NVLink, and nothing orders that flag behind posted stores still in flight
            // (without row_order every rank writes to the same owner at the same time -- the one configuration whose
            // gate failed).
            __threadfence_system();
        }
        if (a.acc) {
            float4 base;
            if (a.acc_mode == 1) {
                const float4 x0 = __ldg(a.x + o);
                base = make_float4(a.alpha * x0.x, a.alpha * x0.y, a.alpha * x0.z, a.alpha * x0.w);
            } else {
                base = __ldcs(a.acc + o);                // running layer mean: streamed
            }
            base.x = fmaf(a.alpha, yv.x, base.x);
            base.y = fmaf(a.alpha, yv.y, base.y);
            base.z = fmaf(a.alpha, yv.z, base.z);
            base.w = fmaf(a.alpha, yv.w, base.w);
            __stcs(a.acc + o, base);
        }
    }
}

template <int LANES, int NV, int TPB = kThreads, bool WIDE = false>
__device__ __forceinline__ void rows_body(const SpmmArgs &a, int64_t block) {
    constexpr int kGroups = TPB / LANES;
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x % LANES;
    int64_t r = block * kGroups + threadIdx.x / LANES;
    if (r >= a.M) return;
    if (a.row_order != nullptr) r = __ldg(a.row_order + r);
    const unsigned mask = group_mask<LANES>(lane);
    const int64_t e_lo = __ldg(a.rowptr + r), e_hi = __ldg(a.rowptr + r + 1);
    if (a.hub_rows != nullptr && e_hi - e_lo > a.hub_threshold) return;  // taken by the hub kernels
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_edges<LANES, NV, WIDE>(a, e_lo, e_hi, gl, mask, acc);
    row_epilogue<LANES, NV>(a, r, gl, acc, e_hi > e_lo);
}

// CTA size and register cap (minimum resident CTAs) are parameters: see launch_spmm.
template <int LANES, int NV, int TPB, int MINB>
__global__ void __launch_bounds__(TPB, MINB) spmm_rows_kernel(const SpmmArgs a) {
    rows_body<LANES, NV, TPB>(a, blockIdx.x);
}

// Sum the partial rows of hub row h in chunk order and apply the row epilogue (one sub-warp).
template <int LANES, int NV>
__device__ __forceinline__ void hub_finish_row(const SpmmArgs &a, int h, int gl) {
    const int64_t r = __ldg(a.hub_rows + h);
    const int64_t c_lo = __ldg(a.hub_chunk_ptr + h), c_hi = __ldg(a.hub_chunk_ptr + h + 1);
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const int q = gl + v * LANES;
        if (q >= a.dq) continue;
        int64_t c = c_lo;
        for (; c + 8 <= c_hi; c += 8) {          // eight independent partial loads in flight, summed in chunk order
            float4 p[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) p[j] = __ldcg(a.hub_partials + (c + j) * a.dq + q);   // written by other SMs: L2
#pragma unroll
            for (int j = 0; j < 8; ++j) { acc[v].x += p[j].x; acc[v].y += p[j].y; acc[v].z += p[j].z; acc[v].w += p[j].w; }
        }
        for (; c < c_hi; ++c) {
            const float4 p = __ldcg(a.hub_partials + c * a.dq + q);
            acc[v].x += p.x; acc[v].y += p.y; acc[v].z += p.z; acc[v].w += p.w;
        }
    }
    row_epilogue<LANES, NV>(a, r, gl, acc);
}

// One CTA per (hub row, chunk): every sub-warp reduces a contiguous slice of the chunk, the
// slices are summed in slice order through shared memory and stored as one partial row.
// Single-launch mode (a.hub_counters): the CTA that completes a row's last outstanding chunk
// sums the row's partials (in chunk order, so the result does not depend on which CTA that is)
// and resets the counter for the next launch.
template <int LANES, int NV>
__device__ __forceinline__ void hub_chunk_body(const SpmmArgs &a, int64_t chunk) {
    constexpr int kGroups = kThreads / LANES;
    __shared__ float4 part[kGroups][LANES * NV];
    __shared__ int is_last;
    const int lane = threadIdx.x & 31;
    const int gl = threadIdx.x % LANES;
    const int g = threadIdx.x / LANES;
    // hub row h with hub_chunk_ptr[h] <= chunk < hub_chunk_ptr[h+1]
    int lo = 0, hi = a.n_hub;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(a.hub_chunk_ptr + mid) <= chunk) lo = mid; else hi = mid;
    }
    const int64_t r = __ldg(a.hub_rows + lo);
    const int64_t c_in_row = chunk - __ldg(a.hub_chunk_ptr + lo);
    const int64_t row_lo = __ldg(a.rowptr + r), row_hi = __ldg(a.rowptr + r + 1);
    const int64_t chunk_edges = hub_chunk_edges(row_hi - row_lo, a.hub_chunk);
    const int64_t c_lo = row_lo + c_in_row * chunk_edges;
    const int64_t c_hi = (c_lo + chunk_edges < row_hi) ? c_lo + chunk_edges : row_hi;
    const int64_t kSlice = (chunk_edges + kGroups - 1) / kGroups;
    int64_t e_lo = c_lo + static_cast<int64_t>(g) * kSlice;
    int64_t e_hi = e_lo + kSlice;
    if (e_lo > c_hi) e_lo = c_hi;
    if (e_hi > c_hi) e_hi = c_hi;
    float4 acc[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) acc[v] = make_float4(0.f, 0.f, 0.f, 0.f);
    gather_edges<LANES, NV, true>(a, e_lo, e_hi, gl, group_mask<LANES>(lane), acc);   // long streams: 8 gathers in flight
#pragma unroll
    for (int v = 0; v < NV; ++v) part[g][gl + v * LANES] = acc[v];
    __syncthreads();
    for (int q = threadIdx.x; q < a.dq; q += kThreads) {
        const int slot = q;  // q = gl + v*LANES by construction of the layout above
        float4 s = part[0][slot];
        for (int gg = 1; gg < kGroups; ++gg) {
            const float4 p = part[gg][slot];
            s.x += p.x; s.y += p.y; s.z += p.z; s.w += p.w;
        }
        a.hub_partials[chunk * a.dq + q] = s;
    }
    if (a.hub_counters == nullptr) return;
    __threadfence();                       // this CTA's partial row is visible device-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {                // ... before its arrival is counted
        const int n_chunks = static_cast<int>(__ldg(a.hub_chunk_ptr + lo + 1) - __ldg(a.hub_chunk_ptr + lo));
        const int prev = atomicAdd(a.hub_counters + lo, 1);
        is_last = (prev == n_chunks - 1);
        if (is_last) a.hub_counters[lo] = 0;
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (g == 0) hub_finish_row<LANES, NV>(a, lo, gl);
}

template <int LANES, int NV>
__global__ void __launch_bounds__(kThreads) spmm_hub_chunk_kernel(const SpmmArgs a) {
    hub_chunk_body<LANES, NV>(a, blockIdx.x);
}

template <int LANES, int NV>
__global__ void __launch_bounds__(kThreads) spmm_hub_finish_kernel(const SpmmArgs a) {
    constexpr int kGroups = kThreads / LANES;
    const int gl = threadIdx.x % LANES;
    const int h = blockIdx.x * kGroups + threadIdx.x / LANES;
    if (h >= a.n_hub) return;
    hub_finish_row<LANES, NV>(a, h, gl);
}

// Single launch for small graphs: the hub-chunk CTAs come first in the grid (they are the long ones), the
// row CTAs follow; rows are finished by the last chunk CTA (hub_counters), so nothing waits on a second launch.
template <int LANES, int NV>
__global__ void __launch_bounds__(kThreads) spmm_fused_kernel(const SpmmArgs a) {
    if (static_cast<int64_t>(blockIdx.x) < a.n_hub_chunks) {
        hub_chunk_body<LANES, NV>(a, blockIdx.x);
        return;
    }
    rows_body<LANES, NV>(a, static_cast<int64_t>(blockIdx.x) - a.n_hub_chunks);
}

template <int LANES, int NV>
int launch_spmm(const SpmmArgs &a, int64_t n_hub_chunks, cudaStream_t st) {
    constexpr int kGroups = kThreads / LANES;
    const int64_t blocks = (a.M + kGroups - 1) / kGroups;
    if (blocks > 0x7fffffffLL) return CHAOREC_E_UNSUPPORTED;
    if (a.hub_counters != nullptr && a.hub_rows != nullptr && n_hub_chunks > 0) {
        if (blocks + n_hub_chunks > 0x7fffffffLL) return CHAOREC_E_UNSUPPORTED;
        spmm_fused_kernel<LANES, NV><<<static_cast<unsigned>(blocks + n_hub_chunks), kThreads, 0, st>>>(a);
        return launch_status();
    } else {
        // Four rows per CTA: a CTA keeps its slots until its longest row is done, so small CTAs waste fewer
        // warp slots on the degree spread inside a block (measured at the 100M-edge config: 15.4 -> 18.1 G
        // edges/s against 256-thread CTAs; 8 gathers in flight or a 32/40-register cap add nothing).
        if constexpr (NV == 1) {
            constexpr int TPB = 4 * LANES, MINB = (LANES >= 32) ? 10 : 20;
            constexpr int G = TPB / LANES;
            const int64_t nb = (a.M + G - 1) / G;
            if (nb > 0x7fffffffLL) return CHAOREC_E_UNSUPPORTED;
            if (nb > 0) spmm_rows_kernel<LANES, NV, TPB, MINB><<<static_cast<unsigned>(nb), TPB, 0, st>>>(a);
        } else {
            if (blocks > 0) spmm_rows_kernel<LANES, NV, kThreads, 1><<<static_cast<unsigned>(blocks), kThreads, 0, st>>>(a);
        }
    }
    if (a.hub_rows != nullptr && a.n_hub > 0 && n_hub_chunks > 0) {
        if (n_hub_chunks > 0x7fffffffLL) return CHAOREC_E_UNSUPPORTED;
        SpmmArgs b = a;
        b.hub_counters = nullptr;          // rows are finished by the finish kernel, not by the last chunk CTA
        spmm_hub_chunk_kernel<LANES, NV><<<static_cast<unsigned>(n_hub_chunks), kThreads, 0, st>>>(b);
        spmm_hub_finish_kernel<LANES, NV><<<(a.n_hub + kGroups - 1) / kGroups, kThreads, 0, st>>>(b);
    }
    return launch_status();
}

// SDDMM on a CSR pattern: out[e] = <a[row(e), :], b[col[e], :]>.  The adjoint of the SpMM with respect to its
// edge values (LATTICE / MICRO learn their item graph through them, Model/LATTICE.py:142-160), and the
// cosine values of a kNN pattern.  One warp per row: the row of `a` stays in registers (d <= 512), every
// edge gathers its row of `b`, the partial dot products are shuffle-reduced.
__global__ void __launch_bounds__(256) sddmm_csr_kernel(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                                                        const float4 *__restrict__ a, const float4 *__restrict__ b,
                                                        float *__restrict__ out, int64_t M, int dq) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * 8 + (threadIdx.x >> 5);
    if (r >= M) return;
    const int lane = threadIdx.x & 31;
    float4 av[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int q = lane + 32 * v;
        av[v] = q < dq ? __ldg(a + r * dq + q) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    const int64_t e_lo = __ldg(rowptr + r), e_hi = __ldg(rowptr + r + 1);
    for (int64_t e = e_lo; e < e_hi; ++e) {
        const int64_t c = __ldg(col + e);
        float s = 0.0f;
#pragma unroll
        for (int v = 0; v < 4; ++v) {
            const int q = lane + 32 * v;
            if (q < dq) {
                const float4 bv = __ldg(b + c * dq + q);
                s = fmaf(av[v].x, bv.x, s); s = fmaf(av[v].y, bv.y, s);
                s = fmaf(av[v].z, bv.z, s); s = fmaf(av[v].w, bv.w, s);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        if (lane == 0) out[e] = s;
    }
}

__global__ void degree_inv_sqrt_kernel(const int64_t *__restrict__ rowptr, int64_t M, float eps,
                                       float *__restrict__ out) {
    const int64_t r = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (r >= M) return;
    const float deg = static_cast<float>(rowptr[r + 1] - rowptr[r]) + eps;
    // IEEE sqrt + division.  deg == 0 (isolated node, eps == 0): the reference's deg.pow(-0.5) is
    // +inf there but no edge ever indexes it; 0 keeps 0 * scale finite in the
    // kernel epilogue, which does multiply the (all-zero) row sum by its row scale.
    out[r] = deg > 0.0f ? 1.0f / sqrtf(deg) : 0.0f;
}

}  // namespace
}  // namespace chaorec

using namespace chaorec;

extern "C" int32_t chaorec_spmm_hub_threshold(void) { return kHubThreshold; }
extern "C" int32_t chaorec_spmm_hub_chunk(void) { return kHubChunk; }

extern "C" int chaorec_spmm_hub_plan(int64_t nnz, int32_t *threshold, int32_t *chunk) {
    if (!threshold || !chunk || nnz < 0) return CHAOREC_E_BADARG;
    // A sub-warp walks its row four gathers at a time (~0.15 us per edge when nothing else hides the
    // latency), so the longest sub-warp row sets the kernel's tail.  Keep that tail near 1/8 of the
    // bandwidth-bound kernel time (nnz / 16 G edges/s): threshold ~ nnz / 16384, a power of two in [64, 1024].
    // Above 256 the spread of row lengths inside a CTA costs more than the extra hub chunks (measured at
    // 100M edges: thresholds 128 / 256 / 1024 / 4096 -> 18.4 / 18.7 / 18.1 / 16.7 G edges/s).
    int t = 64;
    while (t < 256 && static_cast<int64_t>(t) * 2 * 16384 <= nnz) t *= 2;
    *threshold = t;
    *chunk = 4 * t;
    return CHAOREC_OK;
}

extern "C" int64_t chaorec_spmm_hub_row_chunks(int64_t row_nnz, int32_t chunk) {
    if (row_nnz <= 0 || chunk <= 0) return 0;
    const int64_t ce = hub_chunk_edges(row_nnz, chunk);
    return (row_nnz + ce - 1) / ce;
}

extern "C" int chaorec_degree_inv_sqrt(const int64_t *rowptr, int64_t M, float eps, float *out, void *stream) {
    if (!rowptr || !out || M < 0) return CHAOREC_E_BADARG;
    if (M == 0) return CHAOREC_OK;
    const int64_t blocks = (M + 255) / 256;
    degree_inv_sqrt_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(rowptr, M,
                                                                                                         eps, out);
    return launch_status();
}

namespace chaorec {
namespace {
int spmm_entry(const int64_t *rowptr, const int32_t *col, const float *val, const float *row_scale, const float *col_scale,
               const float *x, float *y, float *acc, float alpha, int32_t acc_mode, const int32_t *hub_rows,
               const int64_t *hub_chunk_ptr, int32_t n_hub, int64_t n_hub_chunks, int32_t hub_threshold, int32_t hub_chunk,
               float *hub_partials, int32_t *hub_counters, const int32_t *row_order, int64_t M, int64_t N, int32_t d,
               void *const *peer_y, int32_t n_peers, int64_t peer_block, int32_t peer_slot, void *stream) {
    if (!rowptr || !x || M < 0 || N < 0 || d <= 0) return CHAOREC_E_BADARG;
    if (!y && !acc && n_peers <= 0) return CHAOREC_E_BADARG;
    if (acc_mode != 0 && acc_mode != 1) return CHAOREC_E_BADARG;
    if (acc && acc_mode == 1 && M != N) return CHAOREC_E_BADARG;
    if (d % 4 != 0 || d > 512) return CHAOREC_E_UNSUPPORTED;
    if (!aligned16(x) || (y && !aligned16(y)) || (acc && !aligned16(acc))) return CHAOREC_E_ALIGN;
    if (hub_rows && n_hub > 0 && (!hub_chunk_ptr || !hub_partials || !aligned16(hub_partials)))
        return CHAOREC_E_BADARG;
    if (hub_threshold < 0 || hub_chunk < 0) return CHAOREC_E_BADARG;
    if (n_peers < 0 || n_peers > kMaxPeers) return CHAOREC_E_UNSUPPORTED;
    if (n_peers > 0 && (!peer_y || peer_block <= 0 || peer_slot < 0 || peer_slot >= n_peers || peer_block * n_peers < M))
        return CHAOREC_E_BADARG;
    if (n_peers > 0 && (M > 0x7fffffffLL || peer_block > 0x7fffffffLL)) return CHAOREC_E_UNSUPPORTED;
    if (M == 0) return CHAOREC_OK;
    if (!col) return CHAOREC_E_BADARG;
    SpmmArgs a;
    a.rowptr = rowptr; a.col = col; a.val = val; a.row_scale = row_scale; a.col_scale = col_scale;
    a.x = reinterpret_cast<const float4 *>(x);
    a.y = reinterpret_cast<float4 *>(y);
    a.acc = reinterpret_cast<float4 *>(acc);
    a.alpha = alpha; a.acc_mode = acc_mode;
    a.hub_rows = (hub_rows && n_hub > 0) ? hub_rows : nullptr;
    a.hub_chunk_ptr = hub_chunk_ptr; a.n_hub = a.hub_rows ? n_hub : 0;
    a.hub_partials = reinterpret_cast<float4 *>(hub_partials);
    a.row_order = row_order;
    a.hub_counters = a.hub_rows ? hub_counters : nullptr;
    a.n_hub_chunks = a.hub_rows ? n_hub_chunks : 0;
    a.hub_threshold = hub_threshold > 0 ? hub_threshold : kHubThreshold;
    a.hub_chunk = hub_chunk > 0 ? hub_chunk : kHubChunk;
    a.M = M; a.dq = d / 4;
    a.n_peers = n_peers; a.peer_slot = peer_slot; a.peer_block = peer_block > 0 ? peer_block : 1;
    for (int p = 0; p < kMaxPeers; ++p) {
        a.peer_y[p] = (p < n_peers) ? reinterpret_cast<float4 *>(peer_y[p]) : nullptr;
        if (p < n_peers && (!a.peer_y[p] || !aligned16(a.peer_y[p]))) return CHAOREC_E_ALIGN;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (d == 32) return launch_spmm<8, 1>(a, n_hub_chunks, st);
    if (d == 64) return launch_spmm<16, 1>(a, n_hub_chunks, st);
    if (d <= 128) return launch_spmm<32, 1>(a, n_hub_chunks, st);
    if (d <= 256) return launch_spmm<32, 2>(a, n_hub_chunks, st);
    if (d <= 384) return launch_spmm<32, 3>(a, n_hub_chunks, st);
    return launch_spmm<32, 4>(a, n_hub_chunks, st);
}

// Owner side of the fused exchange: rows [0, n_rows) of this GPU's item block.
//   sum[b] = slots[0][b] + slots[1][b] + ... (rank order: the result does not depend on arrival order)
//   mode 0: acc_blk[b] += alpha * sum[b];  every peer's table row (block_row0 + b) = sum[b]
//   mode 1: acc_blk[b] += alpha * sum[b];  every peer's table row (block_row0 + b) = acc_blk[b]   (last layer)
// One float4 per thread; the G stores of a value go out back to back (G - 1 of them over NVLink).
struct ReduceBcastArgs {
    float4 *slots;
    float4 *dst[kMaxPeers];
    float4 *acc_blk;
    int64_t n_vec;          // n_rows * d / 4
    int64_t slot_stride;    // peer_block * d / 4
    int64_t dst_off;        // block_row0 * d / 4
    int n_peers, mode;
    float alpha;
    const unsigned char *need;   // nullable [n_peers][need_stride]: does peer p read item row r at all?
    int64_t need_stride, row0;
    int dq, self;
};
__global__ void __launch_bounds__(256) reduce_slots_bcast_kernel(const ReduceBcastArgs a) {
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < a.n_vec;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        // all loads of the element first (independent, in flight together), then the sum in rank order, then the stores
        float4 v[kMaxPeers];
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < a.n_peers) v[p] = __ldcs(a.slots + p * a.slot_stride + i);
        float4 c = zero;
        if (a.acc_blk != nullptr) c = a.acc_blk[i];
        float4 s = v[0];
#pragma unroll
        for (int p = 1; p < kMaxPeers; ++p)
            if (p < a.n_peers) { s.x += v[p].x; s.y += v[p].y; s.z += v[p].z; s.w += v[p].w; }
        float4 out = s;
        if (a.acc_blk != nullptr) {
            c.x = fmaf(a.alpha, s.x, c.x); c.y = fmaf(a.alpha, s.y, c.y); c.z = fmaf(a.alpha, s.z, c.z); c.w = fmaf(a.alpha, s.w, c.w);
            a.acc_blk[i] = c;
            if (a.mode == 1) out = c;
        }
        // Intermediate layers: a peer whose users touch no edge of this item never gathers the row -- it is not sent
        // (close to half of the (item, rank) pairs of a Zipf-shaped graph on 8 GPUs).  The last layer's rows are the
        // result and go everywhere.
        const int64_t r = a.row0 + i / a.dq;
#pragma unroll
        for (int p = 0; p < kMaxPeers; ++p)
            if (p < a.n_peers) {
                if (a.need == nullptr || a.mode == 1 || p == a.self || __ldg(a.need + p * a.need_stride + r) != 0) a.dst[p][a.dst_off + i] = out;
                __stcs(a.slots + p * a.slot_stride + i, zero);   // the inbox is zero again for the next layer's (sparse) pushes
            }
    }
    __threadfence_system();     // (the rows stored into the peers' tables have landed when the kernel completes)
}
}  // namespace
}  // namespace chaorec

extern "C" int chaorec_spmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *val,
                                    const float *row_scale, const float *col_scale, const float *x, float *y,
                                    float *acc, float alpha, int32_t acc_mode, const int32_t *hub_rows,
                                    const int64_t *hub_chunk_ptr, int32_t n_hub, int64_t n_hub_chunks,
                                    int32_t hub_threshold, int32_t hub_chunk, float *hub_partials,
                                    int32_t *hub_counters, const int32_t *row_order, int64_t M, int64_t N,
                                    int32_t d, void *stream) {
    return spmm_entry(rowptr, col, val, row_scale, col_scale, x, y, acc, alpha, acc_mode, hub_rows, hub_chunk_ptr, n_hub,
                      n_hub_chunks, hub_threshold, hub_chunk, hub_partials, hub_counters, row_order, M, N, d, nullptr, 0, 0, 0,
                      stream);
}

extern "C" int chaorec_spmm_csr_push_f32(const int64_t *rowptr, const int32_t *col, const float *val,
                                         const float *row_scale, const float *col_scale, const float *x,
                                         const int32_t *hub_rows, const int64_t *hub_chunk_ptr, int32_t n_hub,
                                         int64_t n_hub_chunks, int32_t hub_threshold, int32_t hub_chunk,
                                         float *hub_partials, int32_t *hub_counters, const int32_t *row_order, int64_t M,
                                         int64_t N, int32_t d, void *const *peer_slots, int32_t n_peers,
                                         int64_t peer_block, int32_t my_rank, void *stream) {
    if (n_peers < 1) return CHAOREC_E_BADARG;
    return spmm_entry(rowptr, col, val, row_scale, col_scale, x, nullptr, nullptr, 0.0f, 0, hub_rows, hub_chunk_ptr, n_hub,
                      n_hub_chunks, hub_threshold, hub_chunk, hub_partials, hub_counters, row_order, M, N, d, peer_slots, n_peers,
                      peer_block, my_rank, stream);
}

extern "C" int chaorec_reduce_slots_broadcast_f32(float *slots, int32_t n_peers, int64_t peer_block, int64_t n_rows,
                                                  int32_t d, int64_t block_row0, void *const *peer_tables, float *acc_blk,
                                                  float alpha, int32_t mode, const unsigned char *need, int64_t need_stride,
                                                  int32_t my_rank, void *stream) {
    if (!slots || !peer_tables || n_peers < 1 || n_peers > kMaxPeers || peer_block <= 0 || n_rows < 0 || n_rows > peer_block ||
        d <= 0 || block_row0 < 0 || (mode != 0 && mode != 1) || (mode == 1 && !acc_blk))
        return CHAOREC_E_BADARG;
    if (d % 4 != 0) return CHAOREC_E_UNSUPPORTED;
    if (!aligned16(slots) || (acc_blk && !aligned16(acc_blk))) return CHAOREC_E_ALIGN;
    if (n_rows == 0) return CHAOREC_OK;
    ReduceBcastArgs a{};
    a.slots = reinterpret_cast<float4 *>(slots);
    for (int p = 0; p < kMaxPeers; ++p) {
        a.dst[p] = (p < n_peers) ? reinterpret_cast<float4 *>(peer_tables[p]) : nullptr;
        if (p < n_peers && (!a.dst[p] || !aligned16(a.dst[p]))) return CHAOREC_E_ALIGN;
    }
    a.acc_blk = reinterpret_cast<float4 *>(acc_blk);
    a.n_vec = n_rows * (d / 4);
    a.slot_stride = peer_block * (d / 4);
    a.dst_off = block_row0 * (d / 4);
    a.n_peers = n_peers; a.mode = mode; a.alpha = alpha;
    a.need = need; a.need_stride = need_stride; a.row0 = block_row0; a.dq = d / 4; a.self = my_rank;
    if (need && (need_stride < block_row0 + n_rows || my_rank < 0 || my_rank >= n_peers)) return CHAOREC_E_BADARG;
    int64_t blocks = (a.n_vec + 255) / 256;
    const int64_t cap = static_cast<int64_t>(sm_count()) * 8;      // one resident wave
    if (blocks > cap) blocks = cap;
    reduce_slots_bcast_kernel<<<static_cast<unsigned>(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
    return launch_status();
}

extern "C" int chaorec_sddmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *a, const float *b,
                                     float *out, int64_t M, int32_t d, void *stream) {
    if (!rowptr || !a || !b || M < 0 || d <= 0) return CHAOREC_E_BADARG;
    if (d % 4 != 0 || d > 512) return CHAOREC_E_UNSUPPORTED;
    if (!aligned16(a) || !aligned16(b)) return CHAOREC_E_ALIGN;
    if (M == 0) return CHAOREC_OK;
    if (!col || !out) return CHAOREC_E_BADARG;
    sddmm_csr_kernel<<<static_cast<unsigned>((M + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        rowptr, col, reinterpret_cast<const float4 *>(a), reinterpret_cast<const float4 *>(b), out, M, d / 4);
    return launch_status();
}
