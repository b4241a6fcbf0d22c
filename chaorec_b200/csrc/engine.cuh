// Internal interface between the two top-K engines and the C-ABI entry points.
#pragma once

#include "common.cuh"

namespace chaorec {

// "For every row of A, the K columns of B with the largest canonical dot product."
struct EngineArgs {
    const float *A;   // [Ma, d] fp32
    const float *B;   // [Nb, d] fp32
    int64_t Ma, Nb;
    int d, K;
    const int64_t *mask_rowptr;   // nullable; indexed by the row of A
    const int32_t *mask_col;
    float mask_value;
    int exclude_self;             // skip column (row + self_off)
    int64_t self_off;
    int64_t id_offset;
    int64_t *out_idx;             // [Ma, K]
    float *out_score;             // nullable
    int kp_hint;                  // tensor engine: K' (32, 64 or 128); 0 = choose from K
    const int32_t *row_list;      // nullable: process only rows row_list[0 .. *row_count_dev)
    const unsigned int *row_count_dev;
    // row-list mode: seed_keys[p] (nullable) is a key the p-th listed row is known to reach with K
    // columns (the exact K-th best among its re-scored candidates): the scan then starts with that
    // threshold instead of -inf and lists only the handful of columns that can still matter
    const uint64_t *seed_keys;
};

// exact fp32 CUDA-core engine (rank_topk.cu)
size_t rank_workspace_bytes(int64_t Ma, int d);
size_t rank_listed_workspace_bytes(int d);   // row-list mode (two sibling launches + slice buffer)
int rank_launch_engine(const EngineArgs &e, void *workspace, size_t workspace_bytes, cudaStream_t stm);

// bf16 tcgen05 pre-filter + exact re-score + certification (tc_candidates.cu)
bool tc_supported(int64_t Ma, int64_t Nb, int D, int K);
size_t tc_workspace_bytes(int64_t Ma, int64_t Nb, int D);
int tc_topk_launch(const EngineArgs &e, void *workspace, size_t workspace_bytes, cudaStream_t stm);

// d <= 64: fp16 TS-form tcgen05 pre-filter with per-row-block accumulator slots (tc_eval.cu); tc_topk_launch
// dispatches to it
bool tc_eval_supported(int64_t Ma, int64_t Nb, int D, int K);
size_t tc_eval_workspace_bytes(int64_t Ma, int64_t Nb, int D);
int tc_eval_launch(const EngineArgs &e, void *workspace, size_t workspace_bytes, cudaStream_t stm);
const int32_t *tc_eval_fallback_rows(void *workspace, int64_t Ma, int64_t Nb);

}  // namespace chaorec
