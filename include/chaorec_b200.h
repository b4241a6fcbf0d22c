/*
 * chaorec_b200.h -- C ABI of the B200-native hot path of ChaoRec.
 *
 * The reference (Ricardo-Ping/ChaoRec) is pure Python and has no FFI; its "operator API" for
 * this path is three Python call seams.  Each entry point below names the
 * reference call it replaces (paths relative to the reference checkout).  The Python shims in
 * chaorec_b200/ bind these symbols with ctypes (see INTEGRATION.md for the reference-side
 * stubs a maintainer would add).
 *
 * Conventions (all entry points):
 *   - plain pointers and sizes, no torch types; every pointer is a DEVICE pointer on the
 *     current CUDA device unless the parameter is documented as host memory;
 *   - the caller owns every buffer (outputs and workspace); the library never allocates;
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); no internal sync;
 *   - return 0 on success, a negative CHAOREC_E_* code on bad arguments or launch failure;
 *   - stateless and re-entrant; there is no CPU fallback of any kind.
 */
#ifndef CHAOREC_B200_H_
#define CHAOREC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CHAOREC_ABI_VERSION 3   /* 2: hub plan arguments of the SpMM entry points; 3: fallback-row query, fused multi-GPU exchange entry points, hot-column entry removed */

enum {
    CHAOREC_OK = 0,
    CHAOREC_E_BADARG = -1,       /* null pointer / negative size / inconsistent sizes            */
    CHAOREC_E_UNSUPPORTED = -2,  /* d not a multiple of 4, K out of range, ...                   */
    CHAOREC_E_WORKSPACE = -3,    /* workspace too small                                          */
    CHAOREC_E_LAUNCH = -4,       /* a CUDA call failed (the sticky error is cleared)             */
    CHAOREC_E_ALIGN = -5         /* a pointer that must be 16-byte aligned is not                */
};

int chaorec_abi_version(void);
const char *chaorec_error_string(int code);

/* ------------------------------------------------------------------------------------------
 * (1) Propagation:  y = diag(row_scale) * A * diag(col_scale) * x   (+ fused layer mean)
 * ---------------------------------------------------------------------------------------- */

/*
 * COO -> CSR (stable by row, then column; duplicates are kept and therefore summed by the
 * SpMM exactly as torch.sparse.mm sums an uncoalesced tensor, Model/FREEDOM.py:83,162,174).
 * row/col: int64 [nnz] as stored in torch.sparse_coo_tensor.indices(); val nullable (=> ones).
 * Outputs: rowptr int64 [M+1], out_col int32 [nnz], out_val float [nnz] (ignored if val NULL).
 */
size_t chaorec_coo_to_csr_workspace_bytes(int64_t nnz, int64_t M);
int chaorec_coo_to_csr(const int64_t *row, const int64_t *col, const float *val, int64_t nnz, int64_t M,
                       int64_t N, int64_t *rowptr, int32_t *out_col, float *out_val, void *workspace,
                       size_t workspace_bytes, void *stream);

/*
 * out[r] = (rowptr[r+1]-rowptr[r] + eps)^-1/2 in fp32: deg.pow(-0.5) of Model/LightGCN.py:36-37;
 * with eps = 1e-7 the V3 form of Model/FREEDOM.py:91-96.  An empty row with eps == 0 gets 0 (the
 * reference's +inf is never indexed by an edge; 0 keeps the fused row scaling finite).
 */
int chaorec_degree_inv_sqrt(const int64_t *rowptr, int64_t M, float eps, float *out, void *stream);

/* Rows with more than `threshold` non-zeros are "hub rows"; see chaorec_spmm_csr_f32.  The two
 * no-argument calls return the defaults (used when 0 is passed); chaorec_spmm_hub_plan returns the
 * values recommended for a graph with nnz edges (smaller graphs get a lower threshold so that one long
 * row cannot dominate a kernel that otherwise takes microseconds); chaorec_spmm_hub_row_chunks is
 * the number of chunks a hub row of row_nnz edges is split into for a given `chunk` (host helper:
 * hub_chunk_ptr is the exclusive prefix sum of it over hub_rows). */
int32_t chaorec_spmm_hub_threshold(void);
int32_t chaorec_spmm_hub_chunk(void);
int chaorec_spmm_hub_plan(int64_t nnz, int32_t *threshold, int32_t *chunk);
int64_t chaorec_spmm_hub_row_chunks(int64_t row_nnz, int32_t chunk);

/*
 * CSR SpMM, fp32.  Replaces torch.sparse.mm(adj, x) (Model/FREEDOM.py:168,174 and the 36
 * call sites of SURVEY.md App-A) and LightGCNConv.propagate (Model/LightGCN.py:28-46).
 *
 *   y[r,:] = row_scale[r] * sum_{e in row r} val[e] * col_scale[col[e]] * x[col[e],:]
 *
 * val / row_scale / col_scale are nullable (=> 1).  y is nullable (skip the store when only
 * the layer mean is wanted).  Layer-mean fusion (Model/LightGCN.py:86-93,
 * Model/FREEDOM.py:177-178): if acc != NULL
 *     acc_mode 0:  acc[r,:] += alpha * y[r,:]
 *     acc_mode 1:  acc[r,:]  = alpha * x[r,:] + alpha * y[r,:]     (first layer; needs M == N)
 * Hub rows (nnz > hub_threshold) are split into chunks of at least hub_chunk edges (at most 64
 * chunks per row, chaorec_spmm_hub_row_chunks) reduced through shared memory by whole CTAs: pass
 * the sorted list hub_rows[n_hub], hub_chunk_ptr[n_hub+1] (exclusive prefix of per-row chunk
 * counts) and hub_partials (float [hub_chunk_ptr[n_hub] * d]).  hub_threshold / hub_chunk = 0 mean
 * the defaults.  hub_counters (nullable, int32 [n_hub], zero before the first call; the kernel leaves
 * it zero) selects the single-launch mode meant for small graphs: row CTAs and hub-chunk CTAs share one
 * grid and the CTA that completes a hub row's last chunk sums its partials (in chunk order: results do
 * not depend on scheduling).  The counters are launch state: a hub_counters buffer must not be shared by
 * launches that may run concurrently (different streams); every other argument is read-only or
 * caller-owned output, so the entry point is otherwise re-entrant.  Without it hub rows take two extra launches.  With hub_rows == NULL
 * every row takes the sub-warp path (correct, slow for very long rows).
 * row_order (nullable, int32 [M], a permutation of the rows): the order in which sub-warps take rows.
 * Sorting rows by length (longest first) puts rows of equal length in the same warp and CTA, which
 * removes the idle lanes and slots a mixed block leaves; results do not depend on it.
 * d must be a multiple of 4 and <= 512; x, y, acc 16-byte aligned.
 */
int chaorec_spmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *val, const float *row_scale,
                         const float *col_scale, const float *x, float *y, float *acc, float alpha,
                         int32_t acc_mode, const int32_t *hub_rows, const int64_t *hub_chunk_ptr,
                         int32_t n_hub, int64_t n_hub_chunks, int32_t hub_threshold, int32_t hub_chunk,
                         float *hub_partials, int32_t *hub_counters, const int32_t *row_order, int64_t M, int64_t N,
                         int32_t d, void *stream);

/*
 * Multi-GPU item partials with the exchange fused into the product (one process per GPU, buffers from a symmetric
 * allocation so that every rank holds its peers' addresses).  Same product as chaorec_spmm_csr_f32, but row r is
 * not stored locally: the epilogue of the row stores it over NVLink into the slot buffer of the GPU that owns it,
 *     peer_slots[r / peer_block] + ((my_rank * peer_block + r % peer_block) * d)        (float, [n_peers][peer_block][d])
 * so the reduce-scatter half of the all-reduce of Model/FREEDOM.py:171-178's item table travels while the product
 * runs.  Rows without an edge are NOT sent (the owner keeps its inbox zero: chaorec_reduce_slots_broadcast_f32 writes
 * zeros back as it reads).  peer_slots is a HOST array of n_peers (<= 8) device pointers; peer_block * n_peers >= M < 2^31.
 * The caller separates producer and consumer with a cross-GPU barrier (e.g. torch symmetric-memory barrier).
 */
int chaorec_spmm_csr_push_f32(const int64_t *rowptr, const int32_t *col, const float *val, const float *row_scale,
                              const float *col_scale, const float *x, const int32_t *hub_rows,
                              const int64_t *hub_chunk_ptr, int32_t n_hub, int64_t n_hub_chunks, int32_t hub_threshold,
                              int32_t hub_chunk, float *hub_partials, int32_t *hub_counters, const int32_t *row_order,
                              int64_t M, int64_t N, int32_t d, void *const *peer_slots, int32_t n_peers,
                              int64_t peer_block, int32_t my_rank, void *stream);

/*
 * Owner side of that exchange: for the n_rows (<= peer_block) rows of this GPU's block,
 *     sum[b] = slots[0][b] + slots[1][b] + ... + slots[n_peers-1][b]          (fixed rank order: deterministic);
 *     slots[*][b] = 0                                                          (ready for the next layer's sparse pushes)
 *     acc_blk[b] += alpha * sum[b]                                             (running layer mean of the block; nullable)
 *     row (block_row0 + b) of EVERY peer's table = sum[b] (mode 0) or acc_blk[b] (mode 1, after the last layer)
 * slots: this GPU's [n_peers][peer_block][d] slot buffer; peer_tables: HOST array of n_peers device pointers to the
 * [>= block_row0 + n_rows, d] tables (n_peers - 1 of them are written over NVLink).
 * need (nullable, uint8 [n_peers][need_stride], need_stride >= block_row0 + n_rows): need[p][r] != 0 iff some user of
 * rank p has an edge to item r.  In mode 0 a row is stored only into the tables of the peers that need it (and this
 * rank's own); mode 1 stores every row everywhere.
 */
int chaorec_reduce_slots_broadcast_f32(float *slots, int32_t n_peers, int64_t peer_block, int64_t n_rows, int32_t d,
                                       int64_t block_row0, void *const *peer_tables, float *acc_blk, float alpha,
                                       int32_t mode, const unsigned char *need, int64_t need_stride, int32_t my_rank,
                                       void *stream);

/*
 * SDDMM on a CSR pattern: out[e] = <a[row(e),:], b[col[e],:]> for every stored edge e (fp32, d % 4 == 0,
 * d <= 512; a is [M, d], b is [N, d]).  The gradient of chaorec_spmm_csr_f32 with respect to its edge values
 * (a = dL/dy, b = x) -- the path through which LATTICE / MICRO learn their item-item graph
 * (Model/LATTICE.py:142-160) -- and, with a = b = normalised features, the cosine values of a kNN pattern
 * (build_sim + topk values, Model/LATTICE.py:45-52,67-74).
 */
int chaorec_sddmm_csr_f32(const int64_t *rowptr, const int32_t *col, const float *a, const float *b, float *out,
                          int64_t M, int32_t d, void *stream);

/*
 * BPR step around the propagation (Model/LightGCN.py:97-135): for B triples (users[b], pos[b], neg[b]),
 * rows users[b], item_offset+pos[b], item_offset+neg[b] of emb [N, d]:
 *   loss = -mean log(sigmoid(<u,p> - <u,n>) + 1e-5) + reg_weight * (mean(u^2) + mean(p^2) + mean(n^2))
 * loss_terms float [B] receives the per-triple shares (their sum is the loss), grad float [N*d]
 * must be zero-initialised by the caller and receives d loss / d emb (atomic scatter).  d in {32, 64, 128}.
 */
int chaorec_bpr_loss_grad_f32(const float *emb, const int64_t *users, const int64_t *pos_items,
                              const int64_t *neg_items, int64_t B, int64_t N, int32_t d, int64_t item_offset,
                              float reg_weight, float *loss_terms, float *grad, void *stream);

/* ------------------------------------------------------------------------------------------
 * (2) Full-ranking evaluation
 * ---------------------------------------------------------------------------------------- */

/*
 * Fused score + train-mask + top-K.  Replaces the body of gene_ranklist
 * (Model/LightGCN.py:137-162; batched form Model/MMGCN.py:204-244): the U x I score matrix is
 * never written to memory.
 *   score(u,i) = fp32 fmaf chain over k = 0..d-1 of user_emb[u][k]*item_emb[i][k] (from +0)
 *   masked (u,i) pairs (CSR mask_rowptr int64 [U+1] / mask_col int32, LOCAL item ids, sorted
 *   ascending and unique inside each row) score exactly mask_value (1e-6f; 1e-5f in
 *   Model/GRCN.py:294, Model/MMGCN.py:227, Model/LightGT.py:394)
 *   order: descending score, ties -> ascending item index.
 * out_idx int64 [U*K] = local item id + id_offset (num_user: Model/LightGCN.py:157-159);
 * out_score float [U*K] nullable.  1 <= K <= 64, K <= I, d % 4 == 0.
 * mask_rowptr == NULL => nothing masked.
 */
size_t chaorec_score_mask_topk_workspace_bytes(int64_t U, int64_t I, int32_t d, int32_t K);

/*
 * Two engines compute the same canonical result:
 *   exact   fp32 CUDA cores, every score is the canonical chain (csrc/rank_topk.cu);
 *   tensor  bf16 tcgen05 pre-filter keeps the K' best approximate columns per row, those are
 *           re-scored with the canonical chain and the exact top-K is certified with a rigorous
 *           error bound; uncertified rows are re-run by the exact engine (csrc/tc_candidates.cu).
 * Returns 1 when the tensor engine will be used for this shape (d % 8 == 0, >= 1024 columns; the
 * environment variable CHAOREC_ENGINE=exact forces the exact engine), else 0.  When it is used,
 * the uint32 at byte offset 64 of the workspace holds the number of rows that were re-run by the
 * exact engine once the call has completed.
 */
int chaorec_uses_tensor_engine(int64_t rows, int64_t cols, int32_t d, int32_t K);
/*
 * Which rows were re-run by the exact engine: device pointer (inside `workspace`) of an int32 list whose length is
 * the uint32 at byte offset 64 of the workspace; valid once the chaorec_score_mask_topk_f32 call on that workspace
 * has completed and until the workspace is reused.  NULL when the shape does not use the d <= 64 tensor engine.
 * (Diagnostics: parity gates check every such row against the oracle.)
 */
const int32_t *chaorec_score_mask_topk_fallback_rows(void *workspace, int64_t U, int64_t I, int32_t d, int32_t K);
int chaorec_score_mask_topk_f32(const float *user_emb, const float *item_emb, const int64_t *mask_rowptr,
                                const int32_t *mask_col, float mask_value, int64_t U, int64_t I, int32_t d,
                                int32_t K, int64_t id_offset, int64_t *out_idx, float *out_score,
                                void *workspace, size_t workspace_bytes, void *stream);

/*
 * Hit masks for utils.gene_metrics (utils.py:112-139), one per val_data row: bit r of
 * hit_mask[row] is set iff topk_idx[user*K + r] (user = row_user[row], or row itself when
 * row_user == NULL: utils.py:117-119) occurs in the row's positives
 * pos_col[pos_rowptr[row] .. pos_rowptr[row+1]) (global ids, sorted ascending).  K <= 64.
 */
int chaorec_topk_hits(const int64_t *topk_idx, const int64_t *row_user, const int64_t *pos_rowptr,
                      const int64_t *pos_col, int64_t n_rows, int32_t K, uint64_t *hit_mask, void *stream);

/*
 * HOST function (all pointers are host memory).  The five metrics of metrics.py:13-57 summed
 * sequentially in fp64 over the rows in order and divided by n_rows, exactly as
 * utils.py:116-137 does with python floats.
 *   hit_mask[n_rows]      output of chaorec_topk_hits copied to the host
 *   n_pos_of_row[n_rows]  len(data[1:]) including duplicates (metrics.py:23,31,57)
 *   inv_log[max_k]        1.0/np.log(i+2) computed by numpy on the host (metrics.py:31,34)
 *   out[n_k*5]            precision, recall, ndcg, hit_rate, map for each k
 */
int chaorec_metrics_from_hits(const uint64_t *hit_mask, const int32_t *n_pos_of_row, int64_t n_rows,
                              const int32_t *ks, int32_t n_k, const double *inv_log, double *out);

/* ------------------------------------------------------------------------------------------
 * (3) Modality item-item cosine kNN graph
 * ---------------------------------------------------------------------------------------- */

/* norms[i] = (float)sqrt(sum_k (double)x[i][k]^2) (k ascending, fp64); out[i][:] = x[i][:] / norms[i]
 * (Model/FREEDOM.py:114).  out float [N*D], norms float [N]; both are outputs. */
int chaorec_l2_normalize_f32(const float *x, int64_t N, int32_t D, float *out, float *norms, void *stream);

/*
 * Cosine top-k.  Replaces `torch.topk(build_sim(x), k)` (Model/FREEDOM.py:114-118,
 * Model/LATTICE.py:46-48,68-75, Model/MGCN.py:16-19,58; self excluded:
 * gen_hypergraph_u_i.py:41-46).  Rows [row_lo,row_hi) of the N x N similarity only (item-row
 * sharding across GPUs); out_val float / out_idx int64 are [(row_hi-row_lo)*k], values
 * descending, ties -> ascending column.  The N x N matrix is never materialised.
 * feats: float [N*D] row-major, D % 4 == 0.  1 <= k <= 64.
 */
size_t chaorec_knn_workspace_bytes(int64_t N, int32_t D, int32_t k);
int chaorec_knn_cosine_topk(const float *feats, int64_t N, int32_t D, int32_t k, int exclude_self,
                            int64_t row_lo, int64_t row_hi, float *out_val, int64_t *out_idx, void *workspace,
                            size_t workspace_bytes, void *stream);

/*
 * Negative sampling of the training loader (dataload.py:74-79: draw items uniformly until one is not in
 * user_item_dict[user]).  out[i] = item_offset + c, c the first candidate of draw (first_draw + i) that is
 * not in row users[i] of the train-item CSR (mask_rowptr / mask_col, sorted local ids, the same arrays the
 * evaluation mask uses).  Candidates are splitmix64(seed, draw index, attempt) mapped to [0, num_item) by a
 * multiply-high, so draws are reproducible and independent of batch boundaries; after 256 rejected
 * attempts (a user who interacted with nearly every item) the last candidate is returned.
 */
int chaorec_sample_negatives(const int64_t *users, int64_t n, const int64_t *mask_rowptr, const int32_t *mask_col,
                             int64_t num_item, int64_t item_offset, uint64_t seed, uint64_t first_draw, int64_t *out,
                             void *stream);

/*
 * User-user co-occurrence graph (dualgnn-gen-u-u-matrix.py:20-50,85-103): for every user u in
 * [row_lo, row_hi) the min(k, nnz) users v != u with the largest |items(u) & items(v)|, best first, ties in
 * ascending v.  user_rowptr/user_items: the users' (deduplicated, sorted) local item lists as CSR -- the
 * evaluation mask arrays; item_rowptr/item_users: the transpose (users of every item).  out_idx / out_cnt:
 * int32 [U*k] (row u at u*k; unused slots -1 / 0), out_n int32 [U] = entries written for the row.
 * k <= 256, U <= chaorec_cooccurrence_max_users() (the row's counters live in shared memory), every
 * user with fewer than 65536 items.
 */
int64_t chaorec_cooccurrence_max_users(void);
int chaorec_cooccurrence_topk(const int64_t *user_rowptr, const int32_t *user_items, const int64_t *item_rowptr,
                              const int32_t *item_users, int64_t U, int64_t I, int32_t k, int64_t row_lo, int64_t row_hi,
                              int32_t *out_idx, int32_t *out_cnt, int32_t *out_n, void *stream);

/* ------------------------------------------------------------------------------------------
 * Roofline probes (used by bench.py to measure the denominators it reports against)
 * ---------------------------------------------------------------------------------------- */

/* Runs `iters` dependent-free FFMA per thread on every SM; returns 0 and writes the number of
 * floating-point operations issued to *flops_out (host pointer).  Time it with CUDA events. */
int chaorec_probe_fp32_fma(int64_t iters, float *sink, double *flops_out, void *stream);

/* Reads 32-bit accumulators out of tensor memory with `warps` (1..16) warps per SM and nothing else
 * (tcgen05.ld.32x32b.x32); writes the bytes read to *bytes_out.  The tensor engine's epilogue has to
 * look at every fp32 score once, so this rate bounds rows*cols*4 bytes per pass whatever d is. */
int chaorec_probe_tmem_read(int64_t iters, int32_t warps, uint32_t *sink, double *bytes_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* CHAOREC_B200_H_ */
