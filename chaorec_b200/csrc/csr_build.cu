// COO -> CSR on the device (the step right before K1): stable radix sort of (row, col) keys
// with cub, then a row histogram + exclusive scan for rowptr.  Duplicate (row, col) entries are
// kept: the SpMM sums them, which is what torch.sparse.mm does with an uncoalesced COO tensor
// (Model/FREEDOM.py:83,162 build such tensors).
#include <cub/cub.cuh>

#include "common.cuh"

namespace chaorec {
namespace {

__global__ void pack_keys_kernel(const int64_t *__restrict__ row, const int64_t *__restrict__ col, int64_t nnz,
                                 uint64_t *__restrict__ keys, uint32_t *__restrict__ perm) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    keys[e] = (static_cast<uint64_t>(row[e]) << 32) | static_cast<uint64_t>(static_cast<uint32_t>(col[e]));
    perm[e] = static_cast<uint32_t>(e);
}

__global__ void unpack_kernel(const uint64_t *__restrict__ keys, const uint32_t *__restrict__ perm,
                              const float *__restrict__ val, int64_t nnz, int32_t *__restrict__ out_col,
                              float *__restrict__ out_val, unsigned long long *__restrict__ counts) {
    const int64_t e = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (e >= nnz) return;
    const uint64_t k = keys[e];
    out_col[e] = static_cast<int32_t>(k & 0xFFFFFFFFull);
    if (val && out_val) out_val[e] = val[perm[e]];
    atomicAdd(counts + (k >> 32), 1ull);
}

size_t a256(size_t n) { return (n + 255) & ~static_cast<size_t>(255); }

struct Plan {
    size_t keys_a, keys_b, perm_a, perm_b, counts, cub_tmp, cub_bytes, total;
};

Plan make_plan(int64_t nnz, int64_t M) {
    Plan p{};
    size_t sort_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, static_cast<uint64_t *>(nullptr),
                                    static_cast<uint64_t *>(nullptr), static_cast<uint32_t *>(nullptr),
                                    static_cast<uint32_t *>(nullptr), nnz > 0 ? nnz : 1);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, static_cast<unsigned long long *>(nullptr),
                                  static_cast<int64_t *>(nullptr), M + 1);
    cudaGetLastError();
    p.cub_bytes = sort_bytes > scan_bytes ? sort_bytes : scan_bytes;
    size_t off = 0;
    p.keys_a = off; off += a256(static_cast<size_t>(nnz) * 8);
    p.keys_b = off; off += a256(static_cast<size_t>(nnz) * 8);
    p.perm_a = off; off += a256(static_cast<size_t>(nnz) * 4);
    p.perm_b = off; off += a256(static_cast<size_t>(nnz) * 4);
    p.counts = off; off += a256(static_cast<size_t>(M + 1) * 8);
    p.cub_tmp = off; off += a256(p.cub_bytes);
    p.total = off + 256;
    return p;
}

}  // namespace
}  // namespace chaorec

using namespace chaorec;

extern "C" size_t chaorec_coo_to_csr_workspace_bytes(int64_t nnz, int64_t M) {
    if (nnz < 0 || M < 0) return 0;
    return make_plan(nnz, M).total;
}

extern "C" int chaorec_coo_to_csr(const int64_t *row, const int64_t *col, const float *val, int64_t nnz, int64_t M,
                                  int64_t N, int64_t *rowptr, int32_t *out_col, float *out_val, void *workspace,
                                  size_t workspace_bytes, void *stream) {
    if (!rowptr || nnz < 0 || M < 0 || N < 0 || M > 0x7FFFFFFFLL || N > 0x7FFFFFFFLL) return CHAOREC_E_BADARG;
    if (nnz > 0 && (!row || !col || !out_col)) return CHAOREC_E_BADARG;
    if (val && !out_val) return CHAOREC_E_BADARG;
    if (nnz > 0xFFFFFFF0LL) return CHAOREC_E_UNSUPPORTED;
    const Plan p = make_plan(nnz, M);
    if (!workspace || (reinterpret_cast<uintptr_t>(workspace) & 255u)) return CHAOREC_E_ALIGN;
    if (workspace_bytes < p.total) return CHAOREC_E_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    unsigned char *ws = static_cast<unsigned char *>(workspace);
    uint64_t *keys_a = reinterpret_cast<uint64_t *>(ws + p.keys_a);
    uint64_t *keys_b = reinterpret_cast<uint64_t *>(ws + p.keys_b);
    uint32_t *perm_a = reinterpret_cast<uint32_t *>(ws + p.perm_a);
    uint32_t *perm_b = reinterpret_cast<uint32_t *>(ws + p.perm_b);
    unsigned long long *counts = reinterpret_cast<unsigned long long *>(ws + p.counts);
    void *cub_tmp = ws + p.cub_tmp;
    if (cudaMemsetAsync(counts, 0, static_cast<size_t>(M + 1) * 8, st) != cudaSuccess) {
        cudaGetLastError();
        return CHAOREC_E_LAUNCH;
    }
    if (nnz > 0) {
        const unsigned blocks = static_cast<unsigned>((nnz + 255) / 256);
        pack_keys_kernel<<<blocks, 256, 0, st>>>(row, col, nnz, keys_a, perm_a);
        size_t tmp_bytes = p.cub_bytes;
        int row_bits = 1;
        while ((1LL << row_bits) < M) ++row_bits;
        if (cub::DeviceRadixSort::SortPairs(cub_tmp, tmp_bytes, keys_a, keys_b, perm_a, perm_b, nnz, 0,
                                            32 + row_bits, st) != cudaSuccess) {
            cudaGetLastError();
            return CHAOREC_E_LAUNCH;
        }
        unpack_kernel<<<blocks, 256, 0, st>>>(keys_b, perm_b, val, nnz, out_col, out_val, counts);
    }
    size_t tmp_bytes = p.cub_bytes;
    if (cub::DeviceScan::ExclusiveSum(cub_tmp, tmp_bytes, counts, rowptr, M + 1, st) != cudaSuccess) {
        cudaGetLastError();
        return CHAOREC_E_LAUNCH;
    }
    return launch_status();
}
